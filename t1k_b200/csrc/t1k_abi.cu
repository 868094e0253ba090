// C ABI of the B200 genotyping hot path (include/t1k_b200.h).  Host orchestration in C++, all compute in the
// sm_100a kernels of t1k_kernels.cuh / t1k_pair.cuh / t1k_em.cuh.  There is no CPU fallback: without a CUDA
// device every compute entry point fails with T1K_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/t1k_b200.h"
#include "t1k_comm.hpp"
#include "t1k_em.cuh"
#include "t1k_host.hpp"
#include "t1k_kernels.cuh"
#include "t1k_model.hpp"
#include "t1k_pair.cuh"
#include "t1k_reads.hpp"
#include "t1k_filter.cuh"
#include "t1k_alninfo.cuh"
#include "t1k_ingest.cuh"
#include "t1k_tail.cuh"

using namespace t1k;

namespace {

thread_local std::string g_err;
thread_local bool g_emTrusted = false;
// columns of the EM's matrix handed over by t1k_genotype (see EquivalenceClasses::inPtr): column e = rows[beg[e] .. end[e])
struct EmColumns { const int64_t *beg, *end; const int32_t *rows; size_t nRows; };
thread_local const EmColumns *g_emCols = nullptr;
// the whole matrix already on the device (t1k_tail.cuh): CSR with global offsets (rowPtr[G+1], col), CSC of the rows [g0, g1)
struct EmDevice { const int64_t *rowPtr; const int32_t *col; const int64_t *colBeg, *colEnd; const int32_t *rowIdx; int g0, g1; };
thread_local const EmDevice *g_emDev = nullptr;

int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      char b_[512];                                                                                  \
      snprintf(b_, sizeof(b_), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
      g_err = b_;                                                                                    \
      return T1K_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

#define NK(call)                                                                                     \
  do {                                                                                               \
    int e_ = (call);                                                                                 \
    if (e_ != NCCL_SUCCESS) {                                                                        \
      char b_[512];                                                                                  \
      snprintf(b_, sizeof(b_), "%s:%d %s: %s", __FILE__, __LINE__, #call, nccl().GetErrorString(e_)); \
      g_err = b_;                                                                                    \
      return T1K_ERR_NCCL;                                                                           \
    }                                                                                                \
  } while (0)

// Device workspace pool.  Every hot-path buffer (record store, pairing rows, EM vectors ...) is taken from and returned
// to this per-device free list, so a steady-state call performs no cudaMalloc/cudaFree: those are synchronising driver
// calls whose latency jumps by hundreds of ms when anything else (nvidia-smi polling, another context) holds the
// driver lock.  Best fit within 2x; cached bytes are capped; an allocation failure trims the cache and retries.
struct DevPool {
  struct Block { void *p; size_t bytes; int dev; };
  std::mutex mu;
  std::vector<Block> freeList;
  size_t cached = 0;
  static constexpr size_t kMaxCached = (size_t)96 << 30;

  cudaError_t get(size_t n, void **out, size_t *got) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lk(mu);
      int best = -1;
      for (size_t i = 0; i < freeList.size(); ++i) {
        const Block &b = freeList[i];
        if (b.dev != dev || b.bytes < n || b.bytes > std::max(2 * n, n + ((size_t)64 << 20))) continue;
        if (best < 0 || b.bytes < freeList[best].bytes) best = (int)i;
      }
      if (best >= 0) {
        *out = freeList[best].p; *got = freeList[best].bytes;
        cached -= freeList[best].bytes;
        freeList.erase(freeList.begin() + best);
        return cudaSuccess;
      }
    }
    const size_t want = n < ((size_t)1 << 20) ? ((n + 511) & ~(size_t)511) : ((n + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1));
    cudaError_t e = cudaMalloc(out, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trim(dev, 0);
      e = cudaMalloc(out, want);
    }
    if (e == cudaSuccess) *got = want;
    return e;
  }
  void put(void *p, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    freeList.push_back(Block{p, bytes, dev});
    cached += bytes;
    if (cached > kMaxCached) trim_locked(dev, kMaxCached / 2);
  }
  void trim(int dev, size_t keep) { std::lock_guard<std::mutex> lk(mu); trim_locked(dev, keep); }
  void trim_locked(int dev, size_t keep) {     // frees the largest blocks first
    while (cached > keep) {
      int big = -1;
      for (size_t i = 0; i < freeList.size(); ++i)
        if (freeList[i].dev == dev && (big < 0 || freeList[i].bytes > freeList[big].bytes)) big = (int)i;
      if (big < 0) break;
      cudaFree(freeList[big].p);
      cached -= freeList[big].bytes;
      freeList.erase(freeList.begin() + big);
    }
  }
  size_t cached_bytes() { std::lock_guard<std::mutex> lk(mu); return cached; }
};
DevPool &pool() { static DevPool *p = new DevPool; return *p; }   // leaked on purpose: outlives static destructors

struct DevMem {   // owning device allocation (from the pool)
  void *p = nullptr; size_t bytes = 0;
  DevMem() {}
  DevMem(const DevMem &) = delete;
  DevMem &operator=(const DevMem &) = delete;
  ~DevMem() { release(); }
  void release() { if (p) pool().put(p, bytes); p = nullptr; bytes = 0; }
  cudaError_t alloc(size_t n) {
    if (n == 0) n = 16;
    if (p && bytes >= n && bytes <= std::max(2 * n, n + ((size_t)64 << 20))) return cudaSuccess;   // reuse in place
    release();
    cudaError_t e = pool().get(n, &p, &bytes);
    if (e != cudaSuccess) { p = nullptr; bytes = 0; }
    return e;
  }
  template <class T> T *as() const { return (T *)p; }
  void swap(DevMem &o) { std::swap(p, o.p); std::swap(bytes, o.bytes); }
};

struct PinnedMem {
  void *p = nullptr; size_t bytes = 0;
  ~PinnedMem() { if (p) cudaFreeHost(p); }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMallocHost(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  // grow to at least n bytes keeping the first `keep` bytes
  cudaError_t grow(size_t n, size_t keep) {
    if (n <= bytes) return cudaSuccess;
    size_t want = std::max(n, bytes + bytes / 2);
    void *q = nullptr;
    cudaError_t e = cudaMallocHost(&q, want);
    if (e != cudaSuccess) return e;
    if (p && keep) memcpy(q, p, keep);
    if (p) cudaFreeHost(p);
    p = q; bytes = want;
    return cudaSuccess;
  }
  template <class T> T *as() const { return (T *)p; }
};

// The runtime's "last error" is per host thread and survives until somebody reads it: an error that other code of the
// process (or an unchecked call) left behind must not be mistaken for the failure of the next kernel launch here.  Every
// compute entry point clears it on entry; T1K_DEBUG_STALE=1 reports what was pending.
void stale(const char *tag) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess && getenv("T1K_DEBUG_STALE")) fprintf(stderr, "[t1k] pending CUDA error at %s: %s\n", tag, cudaGetErrorString(e));
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// T1K_TIMING=1: host-side phase timings on stderr (diagnostics only)
struct PhaseTimer {
  bool on; double t;
  PhaseTimer() : on(getenv("T1K_TIMING") != nullptr), t(now_ms()) {}
  void lap(const char *what) { if (on) { double n = now_ms(); fprintf(stderr, "[t1k timing] %-28s %9.2f ms\n", what, n - t); t = n; } }
};

int pick_device(int want, int *out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(T1K_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)"); }
  int dev = want;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
  if (dev >= n) return fail(T1K_ERR_ARG, "device ordinal out of range");
  *out = dev;
  return T1K_OK;
}

}  // namespace

struct T1KRef {
  int device = 0, nSM = 0;
  int32_t nAlleles = 0;
  std::vector<int64_t> offset;       // caller's concatenated layout
  std::vector<u64> wordOff;
  std::vector<int32_t> len;
  size_t paddedBases = 0, covEntries = 0;
  std::vector<u64> covOff;
  DevMem seq2, n2, ex2, dWordOff, dLen, dHasN, dMeta, dSimThr, kinfo, entries, covDiff, covPoint, covFinal, dCovOff;
  RefView R;
  cudaStream_t stream = nullptr;
  cudaStream_t copyStream = nullptr;   // D2H of a chunk's fragment rows while the next chunk's kernels run (t1k_genotype)
  cudaStream_t prepStream = nullptr;   // H2D of the next chunk's raw reads (device-side ingest)
  // launch state of the AssignRead kernels (k_seed / k_deferred / k_passes / k_align), sized on first use
  DevMem candPool, laneScratch, hitBuf, workCtr, errFlag, stats, dq, aq, qCtr;
  u32 candCap = 0, dqCap = 0, aqCap = 0;
  u64 arenaCands = 0;
  int gridBlocks = 0, hitCap = 0, seedCap = 0, scrLen = 0;
  int occ = 6;
  size_t scratchWarps = 0;
  u64 nPostings = 0, nEntries = 0;
  size_t memBudget = 0;      // free device memory right after the reference was uploaded (workspace budget; cudaMemGetInfo is
                             // a slow driver call, so it is not repeated per batch)
  bool covDirty = true;
  PinnedMem pinEntries[2];   // D2H staging of pairing rows (double-buffered by t1k_genotype's chunk pipeline)
  PinnedMem pinSend, pinRecv, pinRecv2; // read-group tables on their way to / from the peers
  PinnedMem pinIds;                     // allele ids of the group entries on their way to the device tail
  // t1k_assign_batch_async: jobs of one reference run in submission order
  std::mutex qMu; std::condition_variable qCv; uint64_t qNext = 0, qServing = 0;
  ~T1KRef() { if (stream) cudaStreamDestroy(stream); if (copyStream) cudaStreamDestroy(copyStream); if (prepStream) cudaStreamDestroy(prepStream); }
};

struct T1KAssignment {
  T1KRef *ref = nullptr;
  int device = 0;          // (kept here: the assignment may be destroyed after its reference)
  u32 nReads = 0;
  DevMem store, storeCtr, readOff, readCnt, readRet, readTop, dMaxCnt;
  u64 storeCap = 0, storeUsed = 0;
  u32 maxCnt = 0;
  unsigned long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float msKernel = 0;
  u32 launches = 0;
};

extern "C" {

const char *t1k_last_error(void) { return g_err.c_str(); }

int t1k_device_count(int *count) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
  if (count) *count = n;
  return T1K_OK;
}

int t1k_ref_create(const T1KRefDesc *d, T1KRef **out) {
  if (!d || !out || d->n_alleles <= 0 || !d->bases || !d->offset || !d->exon_ptr || !d->exon_se) return fail(T1K_ERR_ARG, "t1k_ref_create: bad argument");
  *out = nullptr;
  int dev;
  if (int rc = pick_device(d->device, &dev)) return rc;
  CK(cudaSetDevice(dev));
  if (d->n_alleles >= (1 << 24)) return fail(T1K_ERR_UNSUPPORTED, "more than 2^24 alleles");
  PackedRef P;
  if (!pack_reference(d->n_alleles, d->bases, d->offset, d->exon_ptr, d->exon_se, P))
    return fail(T1K_ERR_ARG, "reference contains a character outside ACGTN");
  for (int i = 0; i < d->n_alleles; ++i)
    if (P.len[i] >= (1 << 22)) return fail(T1K_ERR_UNSUPPORTED, "allele longer than 2^22 bases");
  T1KRef *r = new T1KRef;
  r->device = dev;
  r->nAlleles = d->n_alleles;
  r->offset.assign(d->offset, d->offset + d->n_alleles + 1);
  r->wordOff = P.wordOff; r->len = P.len;
  r->paddedBases = P.totalWords * 32;
  r->covEntries = P.covEntries; r->covOff = P.covOff;
  r->nPostings = P.post.size();
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { delete r; return fail(T1K_ERR_CUDA, cudaGetErrorString(e)); }
  r->nSM = prop.multiProcessorCount;
#define UP(dst, vec)                                                                                        \
  do {                                                                                                      \
    e = r->dst.alloc((vec).size() * sizeof((vec)[0]));                                                      \
    if (e == cudaSuccess) e = cudaMemcpy(r->dst.p, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice); \
    if (e != cudaSuccess) { delete r; return fail(T1K_ERR_CUDA, std::string("t1k_ref_create upload: ") + cudaGetErrorString(e)); } \
  } while (0)
  UP(seq2, P.seq2); UP(n2, P.n2); UP(ex2, P.ex2); UP(dWordOff, P.wordOff); UP(dLen, P.len); UP(dHasN, P.hasN); UP(dMeta, P.meta);
  for (size_t k = 0; k < P.entries.size(); ++k)
    if (P.entries[k].more >= 65535u) { delete r; return fail(T1K_ERR_UNSUPPORTED, "a k-mer occurs at more than 65535 offsets inside one tile of 32 alleles"); }
  r->nEntries = P.entries.size();
  P.entries.resize(P.entries.size() + 4, KmerEntry{0xffffffffu, 0, 0, 0});
  UP(kinfo, P.kinfo); UP(entries, P.entries); UP(dCovOff, P.covOff);
  std::vector<u16> simThr(2 * SIM_DEN);
  sim_threshold_table(d->similarity, simThr.data());
  UP(dSimThr, simThr);
#undef UP
  const size_t covBytes = r->covEntries * sizeof(int32_t);
  e = r->covDiff.alloc(covBytes);
  if (e == cudaSuccess) e = r->covPoint.alloc(covBytes);
  if (e == cudaSuccess) e = r->covFinal.alloc(covBytes);
  if (e == cudaSuccess) e = cudaMemset(r->covDiff.p, 0, covBytes);
  if (e == cudaSuccess) e = cudaMemset(r->covPoint.p, 0, covBytes);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->copyStream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->prepStream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete r; return fail(T1K_ERR_CUDA, std::string("t1k_ref_create: ") + cudaGetErrorString(e)); }
  RefView &R = r->R;
  R.seq2 = r->seq2.as<u64>(); R.n2 = r->n2.as<u64>(); R.ex2 = r->ex2.as<u64>();
  R.wordOff = r->dWordOff.as<u64>(); R.len = r->dLen.as<int32_t>(); R.hasN = r->dHasN.as<u8>(); R.meta = r->dMeta.as<AlleleMeta>(); R.simThr = r->dSimThr.as<u16>();
  R.kstart = nullptr; R.post = nullptr; R.kinfo = r->kinfo.as<KmerInfo>(); R.entries = r->entries.as<KmerEntry>();
  R.covDiff = r->covDiff.as<int32_t>(); R.covPoint = r->covPoint.as<int32_t>(); R.covOff = r->dCovOff.as<u64>();
  R.nAlleles = d->n_alleles; R.sim = d->similarity; R.relax = d->relax_intron;
  {
    size_t freeB = 0, totB = 0;
    if (cudaMemGetInfo(&freeB, &totB) != cudaSuccess) { cudaGetLastError(); freeB = (size_t)8 << 30; }
    r->memBudget = freeB + pool().cached_bytes();
  }
  *out = r;
  return T1K_OK;
}

void t1k_ref_destroy(T1KRef *ref) {
  if (!ref) return;
  cudaSetDevice(ref->device);
  delete ref;
}

int t1k_ref_n_alleles(const T1KRef *ref) { return ref ? ref->nAlleles : 0; }

}  // extern "C"

namespace {

// persistent launch geometry of the AssignRead kernels for reads up to maxLen bases
const void *seed_kernel(int occ) {
  return occ >= 8 ? (const void *)k_seed<8> : occ == 7 ? (const void *)k_seed<7> : occ == 6 ? (const void *)k_seed<6> : occ == 5 ? (const void *)k_seed<5>
       : occ == 4 ? (const void *)k_seed<4> : occ == 3 ? (const void *)k_seed<3> : (const void *)k_seed<2>;
}
int setup_assign_launch(T1KRef *r, int maxLen) {
  int seedCap = std::max(64, (maxLen - KMER + 1 + 31) & ~31);
  // resident blocks per SM k_seed is compiled for (register budget): T1K_ASSIGN_OCC
  int occ = 6;      // measured: 6 blocks/SM (80 registers) beats 4, 5 and 8
  if (const char *env = getenv("T1K_ASSIGN_OCC")) occ = atoi(env);
  if (occ < 2 || occ > 8) occ = 6;
  int hitCap = std::max(1024, 2 * maxLen);     // hits of one allele the hit-list path holds (HBM scratch; a read has <= len - 10 seeds)
  if (const char *env = getenv("T1K_HIT_CAP")) hitCap = std::max(64, atoi(env));
  const int scrLen = maxLen <= 255 ? 255 : MAX_READ_LEN;       // per-lane scratch: two sizes
  if (r->gridBlocks && seedCap <= r->seedCap && occ == r->occ && hitCap <= r->hitCap && scrLen <= r->scrLen) return T1K_OK;
  r->occ = occ;
  const void *kfn = seed_kernel(occ);
  const int rwords = read_words(scrLen);
  const size_t smem = warp_smem_bytes(seedCap, rwords) * WARPS_PER_BLOCK;
  CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSM = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kfn, WARPS_PER_BLOCK * 32, smem));
  if (perSM < 1) return fail(T1K_ERR_UNSUPPORTED, "k_seed does not fit on an SM");
  r->seedCap = seedCap; r->hitCap = hitCap; r->scrLen = scrLen;
  r->gridBlocks = std::min(perSM, occ) * r->nSM;
  const size_t warps = (size_t)r->gridBlocks * WARPS_PER_BLOCK;
  u64 cap = 2ull * (u64)r->nAlleles + 2048;
  if (cap > (1u << 20)) cap = 1u << 20;
  r->candCap = (u32)cap;
  // candidate pool: one arena per k_seed warp; a warp takes read-ends while the arena can hold a worst-case read-end
  u64 arena = std::max<u64>(cap, ((u64)4 << 20) / sizeof(Cand));
  if (const char *env = getenv("T1K_ARENA_CANDS")) arena = std::max<u64>(cap, strtoull(env, nullptr, 10));
  if (warps * arena >= ((u64)1 << 32)) arena = (((u64)1 << 32) - 1) / warps;
  if (arena < cap) return fail(T1K_ERR_UNSUPPORTED, "candidate pool: too many alleles for this launch geometry");
  r->arenaCands = arena;
  r->scratchWarps = warps;
  r->dqCap = 48u << 20; r->aqCap = 48u << 20;      // 1.5 GB + 0.75 GB; a full queue is not an error (the work is done in place)
  if (const char *env = getenv("T1K_QUEUE_ITEMS")) r->dqCap = r->aqCap = (u32)std::max(0l, atol(env));
  CK(r->candPool.alloc(warps * arena * sizeof(Cand)));
  CK(r->laneScratch.alloc(warps * 32 * scr_bytes(scrLen)));
  CK(r->hitBuf.alloc(warps * (size_t)hitCap * 32 * sizeof(u32)));
  CK(r->dq.alloc(std::max<size_t>(1, r->dqCap) * sizeof(DeferItem)));
  CK(r->aq.alloc(std::max<size_t>(1, r->aqCap) * sizeof(AlignItem)));
  CK(r->qCtr.alloc(4 * sizeof(unsigned int)));
  CK(r->workCtr.alloc(sizeof(unsigned int)));
  CK(r->errFlag.alloc(sizeof(int)));
  CK(r->stats.alloc(8 * sizeof(unsigned long long)));
  return T1K_OK;
}

std::string decode_err(int err) {
  std::string s;
  if (err & ERR_BAND) s += " alignment band wider than the supported envelope;";
  if (err & ERR_SCRATCH) s += " chaining scratch overflow;";
  if (err & ERR_EMIT) s += " more than MAX_EMIT seed overlaps for one (read, allele);";
  if (err & ERR_CAND) s += " candidate buffer overflow;";
  if (err & ERR_HITS) s += " more k-mer hits of one read on one allele than the hit-list scratch holds (T1K_HIT_CAP);";
  return s;
}

}  // namespace

namespace {
// Packed read-ends resident on the device: planes / lengths as k_pack_reads (or the ingest kernels) leave them for
// read_words(ref->scrLen) words per plane, weight = number of duplicates.
struct DevReads { const u64 *planes; const u16 *len16; const int32_t *w; u32 n; };

// SeqSet::AssignRead for the batch.  The caller has run setup_assign_launch, reset ref->errFlag / ref->stats and queued the packing
// kernels on ref->stream (their error flags are picked up after the first round).
int assign_core(T1KRef *ref, const DevReads &D, T1KAssignment **out) {
  const u32 n = D.n;
  cudaStream_t st = ref->stream;
  PhaseTimer pt;
  T1KAssignment *a = new T1KAssignment;
  struct Guard { T1KAssignment *a; ~Guard() { delete a; } } guard{a};
  a->ref = ref; a->device = ref->device; a->nReads = n;
  const int rwords = read_words(ref->scrLen);
  CK(a->readOff.alloc((size_t)n * 8)); CK(a->readCnt.alloc((size_t)n * 4)); CK(a->readRet.alloc((size_t)n * 4)); CK(a->readTop.alloc((size_t)n * 4));
  CK(a->storeCtr.alloc(8)); CK(a->dMaxCnt.alloc(4));
  CK(cudaMemsetAsync(a->dMaxCnt.p, 0, 4, st));
  if (n == 0) { guard.a = nullptr; *out = a; return T1K_OK; }
  CK(cudaMemsetAsync(a->storeCtr.p, 0, 8, st));
  // record store: sized from free memory, grown (and only the deferred read-ends re-run) if it fills up
  size_t freeB = ref->memBudget;
  u64 cap = std::max<u64>((u64)n * 6144, 1u << 20);
  const u64 capMax = (u64)(freeB * 0.70) / sizeof(Rec);
  if (cap > capMax) cap = capMax;
  if (const char *envCap = getenv("T1K_STORE_RECORDS")) cap = std::max<u64>(1024, strtoull(envCap, nullptr, 10));
  CK(a->store.alloc(cap * sizeof(Rec)));
  a->storeCap = cap;
  pt.lap("  assign: store alloc");
  AssignParams P;
  P.R = ref->R;
  P.Q.planes = D.planes; P.Q.rwords = rwords; P.Q.maxLen = ref->scrLen; P.Q.len = D.len16; P.Q.weight = D.w; P.Q.workList = nullptr; P.Q.nWork = n;
  P.O.store = a->store.as<Rec>(); P.O.storeCtr = a->storeCtr.as<unsigned long long>(); P.O.storeCap = cap;
  P.O.readOff = a->readOff.as<u64>(); P.O.readCnt = a->readCnt.as<u32>(); P.O.readRet = a->readRet.as<int32_t>(); P.O.readTop = a->readTop.as<u32>();
  P.O.maxCnt = a->dMaxCnt.as<u32>();
  P.O.err = ref->errFlag.as<int>(); P.O.stats = ref->stats.as<unsigned long long>();
  P.candPool = ref->candPool.as<Cand>(); P.arenaCands = ref->arenaCands; P.candCap = ref->candCap; P.laneScratch = ref->laneScratch.as<u8>();
  DevMem dState, dStab;
  CK(dState.alloc((size_t)n * sizeof(ReadState))); CK(dStab.alloc((size_t)n * 2 * 256 * sizeof(u32)));
  P.state = dState.as<ReadState>(); P.stabBuf = dStab.as<u32>();
  P.dq = ref->dq.as<DeferItem>(); P.dqCap = ref->dqCap; P.dqCtr = ref->qCtr.as<unsigned int>();
  P.aq = ref->aq.as<AlignItem>(); P.aqCap = ref->aqCap; P.aqCtr = ref->qCtr.as<unsigned int>() + 1;
  // the DP queue of k_align reuses the DeferItem queue's storage: k_deferred is done with it by then
  P.dpq = (AlignItem *)ref->dq.p; P.dpqCap = (u32)std::min<size_t>((size_t)ref->dqCap * sizeof(DeferItem) / sizeof(AlignItem), 0xffffffffu); P.dpqCtr = ref->qCtr.as<unsigned int>() + 2;
  P.queueMargin = (u32)std::min<size_t>(ref->dqCap / 2, ref->scratchWarps * 2048);
  P.workBegin = 0; P.workEnd = 0;
  { const char *env = getenv("T1K_NO_FAST"); P.noFast = (env && atoi(env) != 0) ? 1 : 0; }
  { const char *env = getenv("T1K_NO_SHARE"); P.noShare = (env && atoi(env) != 0) ? 1 : 0; }
  // the leader list of k_defer_group lives in the AlignItem queue's storage (k_passes fills that queue after k_defer_copy is done)
  const bool canGroup = !P.noShare && (size_t)ref->aqCap * sizeof(AlignItem) >= (size_t)ref->dqCap * sizeof(u32) && ref->dqCap > 0;
  P.lead = canGroup ? (u32 *)ref->aq.p : nullptr; P.leadCtr = ref->qCtr.as<unsigned int>() + 3;
  P.workCtr = ref->workCtr.as<unsigned int>();
  P.hitBuf = ref->hitBuf.as<u32>(); P.hitCap = ref->hitCap; P.seedCap = ref->seedCap;
  DevMem workList;
  std::vector<int32_t> hRet;
  std::vector<u32> todo;            // read-ends waiting for a larger store; empty + first == everything
  bool first = true;
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
  const size_t smem = warp_smem_bytes(ref->seedCap, rwords) * WARPS_PER_BLOCK;
  for (int round = 0;; ++round) {
    if (round > 0) CK(cudaMemsetAsync(ref->errFlag.p, 0, sizeof(int), st));   // round 0 keeps k_pack_reads' flags
    if (first) { P.Q.workList = nullptr; P.Q.nWork = n; }
    else {
      CK(workList.alloc(todo.size() * 4));
      CK(cudaMemcpyAsync(workList.p, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, st));
      P.Q.workList = workList.as<u32>(); P.Q.nWork = (u32)todo.size();
    }
    CK(cudaMemsetAsync(ref->workCtr.p, 0, sizeof(unsigned int), st));
    CK(cudaEventRecord(ev0, st));
    // rounds: k_seed takes read-ends until the warps' candidate arenas are full, then the other three kernels finish them
    const u32 nWork = P.Q.nWork;
    for (u32 done = 0; done < nWork;) {
      CK(cudaMemsetAsync(ref->qCtr.p, 0, 4 * sizeof(unsigned int), st));
      switch (ref->occ) {
        case 8: k_seed<8><<<ref->gridBlocks, WARPS_PER_BLOCK * 32, smem, st>>>(P); break;
        case 7: k_seed<7><<<ref->gridBlocks, WARPS_PER_BLOCK * 32, smem, st>>>(P); break;
        case 6: k_seed<6><<<ref->gridBlocks, WARPS_PER_BLOCK * 32, smem, st>>>(P); break;
        case 5: k_seed<5><<<ref->gridBlocks, WARPS_PER_BLOCK * 32, smem, st>>>(P); break;
        case 4: k_seed<4><<<ref->gridBlocks, WARPS_PER_BLOCK * 32, smem, st>>>(P); break;
        case 3: k_seed<3><<<ref->gridBlocks, WARPS_PER_BLOCK * 32, smem, st>>>(P); break;
        default: k_seed<2><<<ref->gridBlocks, WARPS_PER_BLOCK * 32, smem, st>>>(P); break;
      }
      CK(cudaGetLastError());
      if (P.lead) { k_defer_group<<<ref->nSM * 9, WARPS_PER_BLOCK * 32, 0, st>>>(P); CK(cudaGetLastError()); }     // (no per-lane scratch: full occupancy)
      k_deferred<<<ref->gridBlocks, WARPS_PER_BLOCK * 32, 0, st>>>(P);
      CK(cudaGetLastError());
      if (P.lead) { k_defer_copy<<<ref->nSM * 8, 256, 0, st>>>(P); CK(cudaGetLastError()); a->launches += 2; }
      unsigned int taken = 0;
      CK(cudaMemcpyAsync(&taken, ref->workCtr.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      const u32 roundEnd = std::min<u32>(taken, nWork);
      if (roundEnd <= done) return fail(T1K_ERR_UNSUPPORTED, "t1k_assign_batch: the seeding kernel made no progress");
      P.workBegin = done; P.workEnd = roundEnd;
      k_passes<<<ref->gridBlocks, WARPS_PER_BLOCK * 32, 0, st>>>(P);
      CK(cudaGetLastError());
      k_align<<<ref->gridBlocks, WARPS_PER_BLOCK * 32, 0, st>>>(P);
      CK(cudaGetLastError());
      k_align_dp<<<ref->gridBlocks, WARPS_PER_BLOCK * 32, 0, st>>>(P);
      CK(cudaGetLastError());
      a->launches += 5;
      done = roundEnd;
    }
    CK(cudaEventRecord(ev1, st));
    int err = 0;
    CK(cudaMemcpyAsync(&err, ref->errFlag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, ev0, ev1));
    a->msKernel += ms;
    ref->covDirty = true;
    pt.lap("  assign: kernel round");
    if (err & (ERR_READ_LEN | ERR_READ_CHAR)) return fail(T1K_ERR_ARG, "read contains a character outside ACGTN or is too long");
    if (err & ~ERR_STORE) return fail(T1K_ERR_UNSUPPORTED, "t1k_assign_batch:" + decode_err(err));
    if (!(err & ERR_STORE)) break;
    // deferred read-ends (they added no coverage) wait for a larger store
    hRet.resize(n);
    CK(cudaMemcpy(hRet.data(), a->readRet.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    std::vector<u32> next;
    if (first) { for (u32 i = 0; i < n; ++i) if (hRet[i] == -2) next.push_back(i); }
    else for (size_t k = 0; k < todo.size(); ++k) if (hRet[todo[k]] == -2) next.push_back(todo[k]);
    first = false;
    todo.swap(next);
    {
      u64 newCap = cap * 2;
      if (newCap * sizeof(Rec) > (u64)(freeB * 0.9)) newCap = (u64)(freeB * 0.9) / sizeof(Rec);
      if (newCap <= cap + 1024 || round > 16) return fail(T1K_ERR_UNSUPPORTED, "overlap record store does not fit in device memory; use smaller batches");
      DevMem bigger;
      CK(bigger.alloc(newCap * sizeof(Rec)));
      CK(cudaMemcpyAsync(bigger.p, a->store.p, cap * sizeof(Rec), cudaMemcpyDeviceToDevice, st));
      unsigned long long ctr = cap;   // holes left by failed reservations stay unused
      CK(cudaMemcpyAsync(a->storeCtr.p, &ctr, 8, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      a->store.swap(bigger);
      cap = newCap; a->storeCap = cap;
      P.O.store = a->store.as<Rec>(); P.O.storeCap = cap;
    }
    if (todo.empty()) break;
  }
  unsigned long long used = 0;
  CK(cudaMemcpy(&used, a->storeCtr.p, 8, cudaMemcpyDeviceToHost));
  a->storeUsed = used;
  CK(cudaMemcpy(&a->maxCnt, a->dMaxCnt.p, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(a->stats, ref->stats.p, sizeof(a->stats), cudaMemcpyDeviceToHost));
  if (pt.on) fprintf(stderr, "[t1k timing]   assign: deferred items %llu, distinct within their warp %llu; align items %llu, distinct %llu\n", a->stats[4], a->stats[5], a->stats[6], a->stats[7]);
  pt.lap("  assign: tail");
  guard.a = nullptr;
  *out = a;
  return T1K_OK;
}


}  // namespace

extern "C" {

int t1k_assign_batch(T1KRef *ref, const char *bases, const uint64_t *off, const uint32_t *len, const int32_t *weight,
                     uint32_t n, T1KAssignment **out) {
  if (!ref || !out || (n > 0 && (!bases || !off || !len || !weight))) return fail(T1K_ERR_ARG, "t1k_assign_batch: bad argument");
  *out = nullptr;
  CK(cudaSetDevice(ref->device));
  cudaStream_t st = ref->stream;
  size_t total = 0; int maxLen = KMER;
  for (uint32_t i = 0; i < n; ++i) {
    if (len[i] > T1K_MAX_READ_LEN) return fail(T1K_ERR_ARG, "read longer than T1K_MAX_READ_LEN");
    total = std::max(total, (size_t)(off[i] + len[i]));
    maxLen = std::max(maxLen, (int)len[i]);
  }
  PhaseTimer pt;
  stale("t1k_assign_batch");
  if (int rc = setup_assign_launch(ref, maxLen)) return rc;
  pt.lap("  assign: launch setup");
  DevMem dBases, dOff, dLen, dW, planes, len16;
  CK(dBases.alloc(total)); CK(dOff.alloc((size_t)n * 8)); CK(dLen.alloc((size_t)n * 4)); CK(dW.alloc((size_t)n * 4));
  const int rwords = read_words(ref->scrLen);
  CK(planes.alloc((size_t)n * 4 * rwords * 8)); CK(len16.alloc((size_t)n * 2));
  CK(cudaMemsetAsync(ref->errFlag.p, 0, sizeof(int), st));
  CK(cudaMemsetAsync(ref->stats.p, 0, 8 * sizeof(unsigned long long), st));
  if (n) {
    CK(cudaMemcpyAsync(dBases.p, bases, total, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dOff.p, off, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dLen.p, len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dW.p, weight, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    k_pack_reads<<<(n + 127) / 128, 128, 0, st>>>(dBases.as<char>(), dOff.as<u64>(), dLen.as<u32>(), n, rwords, planes.as<u64>(), len16.as<u16>(),
                                                    ref->errFlag.as<int>());
    CK(cudaGetLastError());
  }
  pt.lap("  assign: input alloc + pack");
  const DevReads D = {planes.as<u64>(), len16.as<u16>(), dW.as<int32_t>(), n};
  return assign_core(ref, D, out);
}

struct T1KAssignJob {
  std::thread worker;
  int rc = T1K_OK;
  std::string err;
  T1KAssignment *result = nullptr;
};

int t1k_assign_batch_async(T1KRef *ref, const char *bases, const uint64_t *off, const uint32_t *len, const int32_t *weight, uint32_t n,
                           T1KAssignJob **job) {
  if (!ref || !job || (n > 0 && (!bases || !off || !len || !weight))) return fail(T1K_ERR_ARG, "t1k_assign_batch_async: bad argument");
  T1KAssignJob *j = new T1KAssignJob;
  uint64_t ticket;
  { std::lock_guard<std::mutex> lk(ref->qMu); ticket = ref->qNext++; }
  j->worker = std::thread([=]() {
    { std::unique_lock<std::mutex> lk(ref->qMu); ref->qCv.wait(lk, [&] { return ref->qServing == ticket; }); }
    j->rc = t1k_assign_batch(ref, bases, off, len, weight, n, &j->result);
    if (j->rc) j->err = g_err;
    { std::lock_guard<std::mutex> lk(ref->qMu); ++ref->qServing; }
    ref->qCv.notify_all();
  });
  *job = j;
  return T1K_OK;
}

int t1k_assign_wait(T1KAssignJob *job, T1KAssignment **out) {
  if (!job || !out) return fail(T1K_ERR_ARG, "t1k_assign_wait: bad argument");
  if (job->worker.joinable()) job->worker.join();
  const int rc = job->rc;
  *out = job->result;
  if (rc) g_err = job->err;
  delete job;
  return rc;
}

int t1k_pinned_alloc(uint64_t bytes, void **p) {
  if (!p) return fail(T1K_ERR_ARG, "t1k_pinned_alloc: bad argument");
  *p = nullptr;
  int dev = 0;
  if (int rc = pick_device(-1, &dev)) return rc;
  CK(cudaMallocHost(p, std::max<uint64_t>(bytes, 16)));
  return T1K_OK;
}
void t1k_pinned_free(void *p) { if (p) cudaFreeHost(p); }

int t1k_assignment_stats(const T1KAssignment *a, T1KAssignStats *out) {
  if (!a || !out) return fail(T1K_ERR_ARG, "t1k_assignment_stats: bad argument");
  out->postings = a->stats[0]; out->candidates = a->stats[1]; out->tiles = a->stats[2]; out->records = a->storeUsed;
  out->ms_kernel = a->msKernel; out->grid_blocks = a->ref->gridBlocks; out->hit_cap = a->ref->hitCap; out->n_sm = a->ref->nSM;
  return T1K_OK;
}

void t1k_assignment_destroy(T1KAssignment *a) {
  if (!a) return;
  cudaSetDevice(a->device);
  delete a;
}

int t1k_assignment_fetch(T1KAssignment *a, uint64_t *row_ptr, int32_t *ret, T1KOverlap *records, uint64_t *total) {
  if (!a || !total) return fail(T1K_ERR_ARG, "t1k_assignment_fetch: bad argument");
  CK(cudaSetDevice(a->ref->device));
  const u32 n = a->nReads;
  std::vector<u64> off(n); std::vector<u32> cnt(n);
  if (n) {
    CK(cudaMemcpy(off.data(), a->readOff.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cnt.data(), a->readCnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  }
  u64 tot = 0;
  for (u32 i = 0; i < n; ++i) tot += cnt[i];
  *total = tot;
  if (row_ptr) { row_ptr[0] = 0; for (u32 i = 0; i < n; ++i) row_ptr[i + 1] = row_ptr[i] + cnt[i]; }
  if (ret && n) CK(cudaMemcpy(ret, a->readRet.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (!records || tot == 0) return T1K_OK;
  const u64 used = std::min(a->storeUsed, a->storeCap);
  std::vector<Rec> st(used);
  CK(cudaMemcpy(st.data(), a->store.p, used * sizeof(Rec), cudaMemcpyDeviceToHost));
  // unpack into the reference's output order (rec_before: list-order key, then store position / extended coordinates)
  std::vector<u32> idx;
  u64 w = 0;
  for (u32 i = 0; i < n; ++i) {
    idx.resize(cnt[i]);
    std::iota(idx.begin(), idx.end(), 0u);
    const Rec *L = st.data() + off[i];
    std::sort(idx.begin(), idx.end(), [L](u32 x, u32 y) { return rec_before(L[x], x, L[y], y); });
    for (u32 k = 0; k < cnt[i]; ++k, ++w) {
      const Rec &r = L[idx[k]];
      T1KOverlap &o = records[w];
      o.seqIdx = r.seqIdx; o.seqStart = r.seqStart; o.seqEnd = r.seqEnd;
      o.readStart = r.readStart; o.readEnd = r.readEnd;
      o.leftClip = r.leftClip; o.rightClip = r.rightClip;
      o.matchCnt = rec_mc(r.mcx); o.strand = rec_strand01(r.mcx) ? 1 : -1;
      o.relaxedMatchCnt = rec_relaxed(r.mcx);
    }
  }
  return T1K_OK;
}

}  // extern "C"

namespace {

__global__ void k_cov_prefix(RefView R, int32_t *covFinal) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= R.nAlleles) return;
  const size_t cb = (size_t)R.covOff[a];        // (the threads of a warp walk the 32 interleaved alleles of a tile: coalesced)
  int run = 0;
  const int n = R.len[a];
  for (int j = 0; j < n; ++j) { const size_t k = cb + (size_t)COV_STRIDE * j; run += R.covDiff[k]; covFinal[k] = run + R.covPoint[k]; }
}

// SeqSet::GetSeqMissingBaseCoverage(a, 0.01) (SeqSet.hpp:2717-2755), one warp per allele: the median of the
// exonic coverages by bisection on the value, then the count below max(1, 1 % of the median).
__global__ void k_missing_coverage(RefView R, const int32_t *covFinal, int32_t *out) {
  const int a = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (a >= R.nAlleles) return;
  const u64 w0 = R.wordOff[a];
  const size_t cb = (size_t)R.covOff[a];
  const int n = R.len[a];
  int nEx = 0, mx = 0;
  for (int j = lane; j < n; j += 32)
    if (base2(R.ex2, w0, j)) { ++nEx; const int c = base2(R.n2, w0, j) ? 0 : covFinal[cb + (size_t)COV_STRIDE * j]; mx = max(mx, c); }
  nEx = warp_sum_i32(nEx); mx = warp_max_i32(mx);
  if (nEx == 0) { if (lane == 0) out[a] = 0; return; }
  const int k = nEx / 2;            // median = element k of the sorted list
  int lo = 0, hi = mx;              // smallest v with count(c <= v) >= k + 1
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    int c = 0;
    for (int j = lane; j < n; j += 32)
      if (base2(R.ex2, w0, j)) { const int v = base2(R.n2, w0, j) ? 0 : covFinal[cb + (size_t)COV_STRIDE * j]; c += v <= mid; }
    c = warp_sum_i32(c);
    if (c >= k + 1) hi = mid; else lo = mid + 1;
  }
  double cutoff = lo * 0.01;
  if (cutoff < 1) cutoff = 1;
  int miss = 0;
  for (int j = lane; j < n; j += 32)
    if (base2(R.ex2, w0, j)) { const int v = base2(R.n2, w0, j) ? 0 : covFinal[cb + (size_t)COV_STRIDE * j]; miss += (double)v < cutoff; }
  miss = warp_sum_i32(miss);
  if (lane == 0) out[a] = miss;
}

int finalize_coverage(T1KRef *ref) {
  if (!ref->covDirty) return T1K_OK;
  stale("coverage");
  k_cov_prefix<<<(ref->nAlleles + 127) / 128, 128, 0, ref->stream>>>(ref->R, ref->covFinal.as<int32_t>());
  CK(cudaGetLastError());
  ref->covDirty = false;
  return T1K_OK;
}

}  // namespace

extern "C" {

int t1k_coverage_fetch(T1KRef *ref, int32_t *out) {
  if (!ref || !out) return fail(T1K_ERR_ARG, "t1k_coverage_fetch: bad argument");
  CK(cudaSetDevice(ref->device));
  if (int rc = finalize_coverage(ref)) return rc;
  std::vector<int32_t> h(ref->covEntries);
  CK(cudaMemcpyAsync(h.data(), ref->covFinal.p, ref->covEntries * 4, cudaMemcpyDeviceToHost, ref->stream));
  CK(cudaStreamSynchronize(ref->stream));
  for (int32_t a = 0; a < ref->nAlleles; ++a)
    for (int32_t j = 0; j < ref->len[a]; ++j) out[ref->offset[a] + j] = h[(size_t)ref->covOff[a] + (size_t)COV_STRIDE * j];
  return T1K_OK;
}

int t1k_coverage_reset(T1KRef *ref) {
  if (!ref) return fail(T1K_ERR_ARG, "t1k_coverage_reset: bad argument");
  CK(cudaSetDevice(ref->device));
  CK(cudaMemsetAsync(ref->covDiff.p, 0, ref->covEntries * 4, ref->stream));
  CK(cudaMemsetAsync(ref->covPoint.p, 0, ref->covEntries * 4, ref->stream));
  CK(cudaStreamSynchronize(ref->stream));
  ref->covDirty = true;
  return T1K_OK;
}

int t1k_missing_coverage(T1KRef *ref, int32_t *out) {
  if (!ref || !out) return fail(T1K_ERR_ARG, "t1k_missing_coverage: bad argument");
  CK(cudaSetDevice(ref->device));
  if (int rc = finalize_coverage(ref)) return rc;
  DevMem d;
  CK(d.alloc((size_t)ref->nAlleles * 4));
  const size_t threads = (size_t)ref->nAlleles * 32;
  k_missing_coverage<<<(unsigned)((threads + 127) / 128), 128, 0, ref->stream>>>(ref->R, ref->covFinal.as<int32_t>(), d.as<int32_t>());
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d.p, (size_t)ref->nAlleles * 4, cudaMemcpyDeviceToHost, ref->stream));
  CK(cudaStreamSynchronize(ref->stream));
  return T1K_OK;
}

}  // extern "C"

namespace {

// Pairing of fragments [0, nFrag) against the resident lists of `a`.  Rows come back in allele order;
// `wantOrder` additionally returns the keys that give the reference's own row order.
struct PairHost {
  std::vector<u64> rowOff;            // [nFrag] first entry of the fragment's row (rows are dense but in no fragment order)
  std::vector<u32> rowCnt;            // [nFrag]
  std::vector<u64> rowHash;           // [2 * nFrag] order-free hash of the row's allele set (when wantHash)
  PinnedMem *pin = nullptr;           // rows land here (pinned: the D2H copy runs at link speed, nothing is zero-filled)
  size_t nEntries = 0;
  HostEntry *entries() const { return pin->as<HostEntry>(); }
  std::vector<u64> ordKey; std::vector<u32> ordIdx;
  std::vector<u8> assigned;           // [nFrag] fragmentAssignment.size() > 0 (before the SetReadAssignments cuts)
  float msKernel = 0;
  u32 launches = 1;
  uint64_t nPairRecords = 0;          // records of both mates' lists summed over the fragments (k_pair's algorithmic reads)
  // asynchronous hand-over (t1k_genotype): the D2H copies run on the copy stream while the next chunk's kernels run; the
  // consumer calls finish() before it reads the rows.  keep[] holds the device buffers the copies read from.
  DevMem keep[4];
  cudaEvent_t copied = nullptr;
  bool pending = false;
  int device = 0;
  ~PairHost() { if (copied) cudaEventDestroy(copied); }
  void unpack_flags() {
    for (size_t i = 0; i < rowCnt.size(); ++i) { assigned[i] = (u8)(rowCnt[i] >> 31); rowCnt[i] &= 0x7fffffffu; }
  }
  cudaError_t finish() {
    if (!pending) return cudaSuccess;
    pending = false;
    cudaSetDevice(device);
    const cudaError_t e = cudaEventSynchronize(copied);
    if (e == cudaSuccess) unpack_flags();
    return e;
  }
};

// devIn: end1 / end2 / hasN are DEVICE arrays the ingest kernels left (valid by construction); else host arrays
int pair_fragments(T1KRef *ref, T1KAssignment *a, const uint32_t *end1, const uint32_t *end2, const uint8_t *hasN, uint32_t nFrag,
                   int maxAssign, bool wantOrder, PairHost &H, bool wantHash = false, bool async = false, bool devIn = false) {
  cudaStream_t st = ref->stream;
  if (H.pending) { if (H.finish() != cudaSuccess) return fail(T1K_ERR_CUDA, "D2H of the previous chunk's fragment rows failed"); }
  for (int k = 0; k < 4; ++k) H.keep[k].release();
  H.rowOff.assign(nFrag, 0); H.rowCnt.assign(nFrag, 0);
  H.rowHash.assign(wantHash ? 2 * (size_t)nFrag : 0, 0);
  H.assigned.assign(nFrag, 0);
  H.nEntries = 0; H.ordKey.clear(); H.ordIdx.clear(); H.nPairRecords = 0; H.msKernel = 0; H.launches = 0;
  if (nFrag == 0) return T1K_OK;
  DevMem dE1, dE2, dN, dRowOff, dRowCnt, dOut, dKey, dIdx, dCtr, dOutCtr, dB0, dStage, dStageKey, dStageIdx, dHash;
  CK(dCtr.alloc(4)); CK(dOutCtr.alloc(8));
  const u32 *pE1 = end1, *pE2 = end2; const u8 *pN = hasN;
  if (devIn) {
    CK(cudaMemsetAsync(dOutCtr.p, 0, 8, st));
    k_pair_records<<<(nFrag + 255) / 256, 256, 0, st>>>(end1, end2, nFrag, a->readCnt.as<u32>(), dOutCtr.as<unsigned long long>());
    CK(cudaGetLastError());
    unsigned long long np = 0;
    CK(cudaMemcpyAsync(&np, dOutCtr.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    H.nPairRecords = np;
  } else {
    {
      std::vector<u32> cnt(a->nReads);
      if (a->nReads) CK(cudaMemcpyAsync(cnt.data(), a->readCnt.p, (size_t)a->nReads * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      for (uint32_t i = 0; i < nFrag; ++i) {
        if (end1[i] < a->nReads) H.nPairRecords += cnt[end1[i]];
        if (end2 && end2[i] < a->nReads) H.nPairRecords += cnt[end2[i]];
      }
    }
    for (uint32_t i = 0; i < nFrag; ++i)
      if (end1[i] >= a->nReads || (end2 && end2[i] >= a->nReads)) return fail(T1K_ERR_ARG, "t1k_pair_batch: read-end index out of range");
    CK(dE1.alloc((size_t)nFrag * 4));
    CK(cudaMemcpyAsync(dE1.p, end1, (size_t)nFrag * 4, cudaMemcpyHostToDevice, st));
    if (end2) { CK(dE2.alloc((size_t)nFrag * 4)); CK(cudaMemcpyAsync(dE2.p, end2, (size_t)nFrag * 4, cudaMemcpyHostToDevice, st)); }
    if (hasN) { CK(dN.alloc(nFrag)); CK(cudaMemcpyAsync(dN.p, hasN, nFrag, cudaMemcpyHostToDevice, st)); }
    pE1 = dE1.as<u32>(); pE2 = end2 ? dE2.as<u32>() : nullptr; pN = hasN ? dN.as<u8>() : nullptr;
  }
  if (wantHash) CK(dHash.alloc((size_t)nFrag * 16));
  CK(dRowOff.alloc((size_t)nFrag * 8)); CK(dRowCnt.alloc((size_t)nFrag * 4));
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
  int pairOcc = 4;
  if (const char *env = getenv("T1K_PAIR_OCC")) pairOcc = atoi(env) == 3 ? 3 : 4;
  PhaseTimer pt;
  // output rows are appended through a device counter; a first guess of the capacity, then (rarely) one exact re-run
  const size_t freeB = ref->memBudget > a->store.bytes ? ref->memBudget - a->store.bytes : ((size_t)1 << 30);
  const size_t perEntry = sizeof(PairEntry) + (wantOrder ? 12 : 0);
  u64 cap = std::max<u64>(1u << 20, (u64)nFrag * 192);
  if (const char *env = getenv("T1K_PAIR_ROWS")) cap = std::max<u64>(64, strtoull(env, nullptr, 10));
  cap = std::min<u64>(cap, (u64)(freeB * 0.6) / perEntry);
  unsigned long long used = 0;
  for (int attempt = 0;; ++attempt) {
    CK(dOut.alloc(cap * sizeof(PairEntry)));
    if (wantOrder) { CK(dKey.alloc(cap * 8)); CK(dIdx.alloc(cap * 4)); }
    CK(cudaMemsetAsync(dCtr.p, 0, 4, st));
    CK(cudaMemsetAsync(dOutCtr.p, 0, 8, st));
    PairParams P;
    P.R = ref->R; P.store = a->store.as<Rec>(); P.readOff = a->readOff.as<u64>(); P.readCnt = a->readCnt.as<u32>(); P.readTop = a->readTop.as<u32>();
    P.end1 = pE1; P.end2 = pE2; P.hasN = pN;
    P.fragBase = 0; P.nFrag = nFrag; P.maxAssign = maxAssign;
    P.out = dOut.as<PairEntry>(); P.outCap = cap; P.outCtr = dOutCtr.as<unsigned long long>(); P.rowOff = dRowOff.as<u64>();
    P.ordKey = wantOrder ? dKey.as<u64>() : nullptr; P.ordIdx = wantOrder ? dIdx.as<u32>() : nullptr;
    P.rowCnt = dRowCnt.as<u32>(); P.rowHash = wantHash ? dHash.as<u64>() : nullptr; P.workCtr = dCtr.as<unsigned int>();
    const int blocks = std::max(1, std::min<int>((int)((nFrag + 3) / 4), ref->nSM * 8));
    // per-warp scratch: where each allele run of the first mate's list starts in the second mate's list
    const size_t b0Stride = ((size_t)a->maxCnt + 32) & ~(size_t)31;
    // ... and the staging row: a row longer than -n is cut, so -n entries suffice
    const size_t stageCap = (maxAssign > 0 && (size_t)maxAssign < b0Stride) ? (size_t)maxAssign : b0Stride;
    if (attempt == 0) {
      CK(dB0.alloc((size_t)blocks * 4 * b0Stride * 2 * 4));      // per warp: b0[b0Stride] + keys[b0Stride]
      CK(dStage.alloc((size_t)blocks * 4 * stageCap * sizeof(PairEntry)));
      if (wantOrder) { CK(dStageKey.alloc((size_t)blocks * 4 * stageCap * 8)); CK(dStageIdx.alloc((size_t)blocks * 4 * stageCap * 4)); }
    }
    P.b0 = dB0.as<u32>(); P.b0Stride = (u32)b0Stride;
    P.stage = dStage.as<PairEntry>(); P.stageCap = (u32)stageCap;
    P.stageKey = wantOrder ? dStageKey.as<u64>() : nullptr; P.stageIdx = wantOrder ? dStageIdx.as<u32>() : nullptr;
    CK(cudaEventRecord(ev0, st));
    if (pairOcc == 3) k_pair<3><<<blocks, 128, 0, st>>>(P); else k_pair<4><<<blocks, 128, 0, st>>>(P);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev1, st));
    CK(cudaMemcpyAsync(&used, dOutCtr.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, ev0, ev1)); H.msKernel += ms; H.launches += 1;
    pt.lap("  pair: alloc + kernel");
    if (used <= cap) break;
    if (attempt >= 2 || used * perEntry > (u64)(freeB * 0.9)) return fail(T1K_ERR_UNSUPPORTED, "fragment rows do not fit in device memory; use smaller batches");
    cap = used;
  }
  H.nEntries = used;
  if (async && !wantOrder) {
    // the kernel is done (its counter has been read): the rows travel on the copy stream and the caller goes on
    cudaStream_t cs = ref->copyStream;
    if (used > 0) CK(H.pin->grow(used * sizeof(HostEntry), 0));
    CK(cudaMemcpyAsync(H.rowOff.data(), dRowOff.p, (size_t)nFrag * 8, cudaMemcpyDeviceToHost, cs));
    CK(cudaMemcpyAsync(H.rowCnt.data(), dRowCnt.p, (size_t)nFrag * 4, cudaMemcpyDeviceToHost, cs));
    if (wantHash) CK(cudaMemcpyAsync(H.rowHash.data(), dHash.p, (size_t)nFrag * 16, cudaMemcpyDeviceToHost, cs));
    if (used > 0) CK(cudaMemcpyAsync(H.entries(), dOut.p, used * sizeof(PairEntry), cudaMemcpyDeviceToHost, cs));
    if (!H.copied) CK(cudaEventCreateWithFlags(&H.copied, cudaEventDisableTiming));
    CK(cudaEventRecord(H.copied, cs));
    H.keep[0].swap(dRowOff); H.keep[1].swap(dRowCnt); H.keep[2].swap(dHash); H.keep[3].swap(dOut);
    H.pending = true; H.device = ref->device;
    return T1K_OK;
  }
  CK(cudaMemcpyAsync(H.rowOff.data(), dRowOff.p, (size_t)nFrag * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(H.rowCnt.data(), dRowCnt.p, (size_t)nFrag * 4, cudaMemcpyDeviceToHost, st));
  if (wantHash) CK(cudaMemcpyAsync(H.rowHash.data(), dHash.p, (size_t)nFrag * 16, cudaMemcpyDeviceToHost, st));
  if (used > 0) {
    CK(H.pin->grow(used * sizeof(HostEntry), 0));
    CK(cudaMemcpyAsync(H.entries(), dOut.p, used * sizeof(PairEntry), cudaMemcpyDeviceToHost, st));
    if (wantOrder) {
      H.ordKey.resize(used); H.ordIdx.resize(used);
      CK(cudaMemcpyAsync(H.ordKey.data(), dKey.p, used * 8, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(H.ordIdx.data(), dIdx.p, used * 4, cudaMemcpyDeviceToHost, st));
    }
  }
  CK(cudaStreamSynchronize(st));
  pt.lap("  pair: D2H");
  H.unpack_flags();
  return T1K_OK;
}

}  // namespace

extern "C" {

int t1k_pair_batch(T1KRef *ref, T1KAssignment *a, const uint32_t *end1, const uint32_t *end2, const uint8_t *has_n,
                   uint32_t n_frag, int32_t max_assign, uint64_t **row_ptr, T1KReadAssignment **entries,
                   uint8_t *fragment_assigned) {
  if (!ref || !a || !row_ptr || !entries || (n_frag > 0 && !end1)) return fail(T1K_ERR_ARG, "t1k_pair_batch: bad argument");
  static_assert(sizeof(T1KReadAssignment) == sizeof(HostEntry) && sizeof(HostEntry) == sizeof(PairEntry), "layout");
  CK(cudaSetDevice(ref->device));
  stale("t1k_pair_batch");
  PairHost H;
  H.pin = &ref->pinEntries[0];
  if (int rc = pair_fragments(ref, a, end1, end2, has_n, n_frag, max_assign, true, H)) return rc;
  uint64_t *rp = (uint64_t *)malloc(((size_t)n_frag + 1) * 8);
  T1KReadAssignment *en = (T1KReadAssignment *)malloc(std::max<size_t>(1, H.nEntries) * sizeof(T1KReadAssignment));
  if (!rp || !en) { free(rp); free(en); return fail(T1K_ERR_ARG, "out of host memory"); }
  rp[0] = 0;
  for (uint32_t f = 0; f < n_frag; ++f) rp[f + 1] = rp[f] + H.rowCnt[f];
  // the reference's row order: by the list position of each allele's first candidate (SeqSet.hpp:2440-2455)
  std::vector<u32> idx;
  for (uint32_t f = 0; f < n_frag; ++f) {
    const u64 src = H.rowOff[f], n = H.rowCnt[f], dst = rp[f];
    idx.resize(n);
    std::iota(idx.begin(), idx.end(), 0u);
    const u64 *k = H.ordKey.data() + src; const u32 *ki = H.ordIdx.data() + src;
    std::sort(idx.begin(), idx.end(), [k, ki](u32 x, u32 y) { return k[x] != k[y] ? k[x] < k[y] : ki[x] < ki[y]; });
    for (u64 j = 0; j < n; ++j) memcpy(&en[dst + j], H.entries() + src + idx[j], sizeof(HostEntry));
  }
  if (fragment_assigned && n_frag) memcpy(fragment_assigned, H.assigned.data(), n_frag);
  *row_ptr = rp; *entries = en;
  return T1K_OK;
}

void t1k_free(void *p) { free(p); }

int t1k_em_run(const T1KEmProblem *p, T1KEmResult *r, int32_t device) {
  if (!p || !r || !r->x || !r->ec_read_count || p->n_ec <= 0 || p->n_groups < 0 || !p->row_ptr || !p->ec_len || !p->x0)
    return fail(T1K_ERR_ARG, "t1k_em_run: bad argument");
  int dev;
  if (int rc = pick_device(device, &dev)) return rc;
  CK(cudaSetDevice(dev));
  stale("t1k_em_run");
  const int G = p->n_groups, E = p->n_ec;
  const int64_t nnz = p->row_ptr[G];
  PhaseTimer pt;
  const EmDevice *emd = g_emTrusted ? g_emDev : nullptr;
  if (!g_emTrusted)        // (t1k_genotype builds the matrix itself)
    for (int64_t k = 0; k < nnz; ++k) if (p->col[k] < 0 || p->col[k] >= E) return fail(T1K_ERR_ARG, "t1k_em_run: column index out of range");
  // read-sharded E-step: this rank's contiguous row range [g0, g1)
  T1KComm *comm = (p->comm && p->comm->world > 1) ? p->comm : nullptr;
  int g0 = 0, g1 = G;
  if (comm) {
    if (!nccl().load()) return fail(T1K_ERR_NCCL, nccl().error);
    std::vector<int32_t> bounds((size_t)comm->world + 1);
    partition_rows(p->row_ptr, G, comm->world, bounds.data());
    g0 = bounds[comm->rank]; g1 = bounds[comm->rank + 1];
    if (emd && (emd->g0 != g0 || emd->g1 != g1)) return fail(T1K_ERR_ARG, "t1k_em_run: device matrix built for another row range");
  }
  const int Gl = g1 - g0;
  const int64_t k0 = p->row_ptr[g0], nnzL = p->row_ptr[g1] - k0;
  std::vector<int64_t> rowPtrL((size_t)(emd ? 0 : Gl + 1));
  if (!emd) for (int g = 0; g <= Gl; ++g) rowPtrL[g] = p->row_ptr[g0 + g] - k0;
  // CSC of the local rows with ascending group order inside every column => fixed summation order
  std::vector<int64_t> colPtr;
  std::vector<int32_t> rowIdx;
  const EmColumns *cols = (!comm && g_emTrusted && !emd) ? g_emCols : nullptr;
  if (!cols && !emd) {
    int hostThreads = (int)std::thread::hardware_concurrency() / (comm ? comm->world : 1);
    if (const char *env = getenv("T1K_HOST_THREADS")) hostThreads = atoi(env);
    const int32_t *colp = p->col;
    transpose_csr(rowPtrL.data(), Gl, [colp, k0](int64_t k) { return colp[k0 + k]; }, E, std::max(1, std::min(8, hostThreads)), colPtr, rowIdx);
    for (size_t k = 0; k < (size_t)nnzL; ++k) rowIdx[k] += g0;       // global group ids
  }
  pt.lap("  em: validate + CSC");
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StGuard { cudaStream_t s; ~StGuard() { cudaStreamDestroy(s); } } sg{st};
  DevMem dRowPtr, dCol, dColPtr, dColEnd, dRowIdx, dCount, dLen, dPsum, dRc, dX0, dX1, dX2, dX3, dDiff, dTmpA, dTmpB;
  // reference-order sums only where they can give the reference's bits: one GPU; a row-sharded run adds the per-rank sums in
  // another order anyway (documented tolerance 1e-5), so it takes the tree reductions
  const bool fast = p->fast_sums != 0 || comm != nullptr;
  CK(dTmpA.alloc((size_t)E * 8)); CK(dTmpB.alloc((size_t)E * 8));
  if (!emd) { CK(dRowPtr.alloc(((size_t)Gl + 1) * 8)); CK(dCol.alloc((size_t)nnzL * 4)); CK(dColPtr.alloc(((size_t)E + 1) * 8)); CK(dRowIdx.alloc((cols ? cols->nRows : (size_t)nnzL) * 4)); }
  if (cols) CK(dColEnd.alloc((size_t)E * 8));
  CK(dCount.alloc((size_t)G * 8)); CK(dLen.alloc((size_t)E * 4)); CK(dPsum.alloc((size_t)G * 8)); CK(dRc.alloc((size_t)E * 8));
  CK(dX0.alloc((size_t)E * 8)); CK(dX1.alloc((size_t)E * 8)); CK(dX2.alloc((size_t)E * 8)); CK(dX3.alloc((size_t)E * 8)); CK(dDiff.alloc(8));
  if (!emd) CK(cudaMemcpyAsync(dRowPtr.p, rowPtrL.data(), ((size_t)Gl + 1) * 8, cudaMemcpyHostToDevice, st));
  if (!emd && nnzL) CK(cudaMemcpyAsync(dCol.p, p->col + k0, (size_t)nnzL * 4, cudaMemcpyHostToDevice, st));
  if (emd) {
  } else if (cols) {
    CK(cudaMemcpyAsync(dColPtr.p, cols->beg, (size_t)E * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dColEnd.p, cols->end, (size_t)E * 8, cudaMemcpyHostToDevice, st));
    if (cols->nRows) CK(cudaMemcpyAsync(dRowIdx.p, cols->rows, cols->nRows * 4, cudaMemcpyHostToDevice, st));
  } else {
    CK(cudaMemcpyAsync(dColPtr.p, colPtr.data(), ((size_t)E + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nnzL) CK(cudaMemcpyAsync(dRowIdx.p, rowIdx.data(), (size_t)nnzL * 4, cudaMemcpyHostToDevice, st));
  }
  const int64_t *dBeg = emd ? emd->colBeg : dColPtr.as<int64_t>(), *dEnd = emd ? emd->colEnd : cols ? dColEnd.as<int64_t>() : dColPtr.as<int64_t>() + 1;
  const int64_t *dRowPtrP = emd ? emd->rowPtr + g0 : dRowPtr.as<int64_t>();      // (global offsets into the whole col array)
  const int32_t *dColP = emd ? emd->col : dCol.as<int32_t>();
  const int32_t *dRowIdxP = emd ? emd->rowIdx : dRowIdx.as<int32_t>();
  if (G) CK(cudaMemcpyAsync(dCount.p, p->count, (size_t)G * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dLen.p, p->ec_len, (size_t)E * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dX0.p, p->x0, (size_t)E * 8, cudaMemcpyHostToDevice, st));
  const unsigned gRow = (unsigned)(((size_t)Gl * 32 + 255) / 256), gCol = (unsigned)(((size_t)E * 32 + 255) / 256);
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
  CK(cudaEventRecord(ev0, st));
  CK(cudaStreamSynchronize(st));
  pt.lap("  em: alloc + upload");
  uint64_t launches = 0;
  auto em_update = [&](const double *xin, double *xout) -> int {   // Genotyper::EMupdate
    double *psumL = dPsum.as<double>() + g0;    // psum / count stay indexed by the global group id
    if (fast) {
      if (Gl) k_em_rowsum<<<gRow, 256, 0, st>>>(Gl, dRowPtrP, dColP, xin, psumL);
      k_em_colsum<<<gCol, 256, 0, st>>>(E, dBeg, dEnd, dRowIdxP, dCount.as<double>(), dPsum.as<double>(), xin, dRc.as<double>());
    } else {
      if (Gl) k_em_rowsum_seq<<<(Gl + 127) / 128, 128, 0, st>>>(Gl, dRowPtrP, dColP, xin, psumL);
      k_em_colsum_seq<<<(unsigned)(((size_t)E * 32 + 255) / 256), 256, 0, st>>>(E, dBeg, dEnd, dRowIdxP, dCount.as<double>(), dPsum.as<double>(), xin, dRc.as<double>());
    }
    CK(cudaGetLastError());
    // the one exchange of the EM: per-EC expected read counts summed over the row shards (NVLink all-reduce)
    if (comm) NK(nccl().AllReduce(dRc.p, dRc.p, (size_t)E, NCCL_FLOAT64, NCCL_SUM, comm->comm, st));
    if (fast) k_em_mstep<<<1, 1024, 0, st>>>(E, dRc.as<double>(), dLen.as<int32_t>(), xout);
    else k_em_mstep_seq<<<1, 1024, 0, st>>>(E, dRc.as<double>(), dLen.as<int32_t>(), dTmpA.as<double>(), xout);
    CK(cudaGetLastError());
    launches += Gl ? 3 : 2;
    return T1K_OK;
  };
  const int maxIter = 1000;
  int ret = 0;
  std::vector<double> hRc(E), hX(E);
  const bool mask = p->n_alleles > 0 && p->ec_allele_ptr && p->ec_alleles && p->allele_major && p->allele_gene;
  // One SQUAREM iteration (Genotyper.hpp:1236-1287: EMupdate x0->x1, x1->x2, extrapolation -> x3, EMupdate x3->x1, advance) is the
  // same eleven kernels (+ three all-reduces) on the same buffers every time: captured once into a CUDA graph and replayed.
  // The host only reads diffSum (8 bytes) after each replay — the reference's stop test decides iteration by iteration.
  auto iteration = [&]() -> int {
    if (int rc = em_update(dX0.as<double>(), dX1.as<double>())) return rc;
    if (int rc = em_update(dX1.as<double>(), dX2.as<double>())) return rc;
    if (fast) k_em_squarem<<<1, 1024, 0, st>>>(E, dX0.as<double>(), dX1.as<double>(), dX2.as<double>(), p->min_squarem_alpha, dX3.as<double>());
    else k_em_squarem_seq<<<1, 1024, 0, st>>>(E, dX0.as<double>(), dX1.as<double>(), dX2.as<double>(), p->min_squarem_alpha, dTmpA.as<double>(), dTmpB.as<double>(), dX3.as<double>());
    if (int rc = em_update(dX3.as<double>(), dX1.as<double>())) return rc;
    if (fast) k_em_advance<<<1, 1024, 0, st>>>(E, dX0.as<double>(), dX1.as<double>(), dDiff.as<double>());
    else k_em_advance_seq<<<1, 1024, 0, st>>>(E, dX0.as<double>(), dX1.as<double>(), dTmpA.as<double>(), dDiff.as<double>());
    CK(cudaGetLastError());
    launches += 2;
    return T1K_OK;
  };
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graphExec = nullptr;
  struct GraphGuard { cudaGraph_t &g; cudaGraphExec_t &e; ~GraphGuard() { if (e) cudaGraphExecDestroy(e); if (g) cudaGraphDestroy(g); } } gg{graph, graphExec};
  bool useGraph = getenv("T1K_EM_NO_GRAPH") == nullptr;
  uint64_t launchesPerIter = 0;
  if (useGraph) {
    const uint64_t before = launches;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); useGraph = false; }
    else {
      const int rcCap = iteration();
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rcCap != T1K_OK || ce != cudaSuccess || cudaGraphInstantiate(&graphExec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        useGraph = false;
        if (rcCap != T1K_OK && comm) return rcCap;       // (a failed collective inside the capture is not recoverable)
      }
    }
    launchesPerIter = launches - before;
    launches = before;
  }
  for (int t = 0; t < maxIter; ++t) {     // Genotyper.hpp:1234-1314
    ++ret;
    if (useGraph) { CK(cudaGraphLaunch(graphExec, st)); launches += launchesPerIter; }
    else if (int rc = iteration()) return rc;
    double diff = 0;
    CK(cudaMemcpyAsync(&diff, dDiff.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (diff < 1e-5 && t < maxIter - 2) t = maxIter - 2;
    if (t > 0 && t % 10 == 0 && mask) {
      // the every-10 low-abundance mask (Genotyper.hpp:1292-1313) runs once or twice per sample, on the host in the reference's
      // own summation order, on a 8 E-byte copy of ecReadCount
      CK(cudaMemcpyAsync(hRc.data(), dRc.p, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      em_mask(hRc.data(), p->ec_len, p->ec_allele_ptr, p->ec_alleles, E, p->n_alleles, p->allele_major, p->allele_gene, p->n_major,
              p->n_gene, p->filter_frac, hX.data());
      CK(cudaMemcpyAsync(dX0.p, hX.data(), (size_t)E * 8, cudaMemcpyHostToDevice, st));
    }
  }
  CK(cudaMemcpyAsync(r->x, dX0.p, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(r->ec_read_count, dRc.p, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ev1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&r->ms_kernel, ev0, ev1));
  pt.lap("  em: iterations");
  r->n_launches = launches;
  r->iterations = ret;
  return T1K_OK;
}

struct T1KGroups { ReadGroups G; };

int t1k_comm_unique_id(uint8_t *id) {
  if (!id) return fail(T1K_ERR_ARG, "t1k_comm_unique_id: bad argument");
  if (!nccl().load()) return fail(T1K_ERR_NCCL, nccl().error);
  ncclUniqueId u;
  NK(nccl().GetUniqueId(&u));
  static_assert(sizeof(u) == T1K_UNIQUE_ID_BYTES, "ncclUniqueId size");
  memcpy(id, &u, sizeof(u));
  return T1K_OK;
}

int t1k_comm_create(const uint8_t *id, int32_t rank, int32_t world, int32_t device, T1KComm **out) {
  if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(T1K_ERR_ARG, "t1k_comm_create: bad argument");
  *out = nullptr;
  int dev;
  if (int rc = pick_device(device, &dev)) return rc;
  if (!nccl().load()) return fail(T1K_ERR_NCCL, nccl().error);
  CK(cudaSetDevice(dev));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  T1KComm *c = new T1KComm;
  c->rank = rank; c->world = world; c->device = dev;
  int e = nccl().CommInitRank(&c->comm, world, u, rank);
  if (e != NCCL_SUCCESS) { delete c; return fail(T1K_ERR_NCCL, std::string("ncclCommInitRank: ") + nccl().GetErrorString(e)); }
  *out = c;
  return T1K_OK;
}

void t1k_comm_destroy(T1KComm *c) {
  if (!c) return;
  if (c->comm) { cudaSetDevice(c->device); nccl().CommDestroy(c->comm); }
  delete c;
}

int t1k_coverage_allreduce(T1KRef *ref, T1KComm *comm) {
  if (!ref || !comm) return fail(T1K_ERR_ARG, "t1k_coverage_allreduce: bad argument");
  if (comm->world == 1) return T1K_OK;
  CK(cudaSetDevice(ref->device));
  // coverage = prefix(covDiff) + covPoint is linear in both arrays: sum them where they lie
  NK(nccl().AllReduce(ref->covDiff.p, ref->covDiff.p, ref->covEntries, NCCL_INT32, NCCL_SUM, comm->comm, ref->stream));
  NK(nccl().AllReduce(ref->covPoint.p, ref->covPoint.p, ref->covEntries, NCCL_INT32, NCCL_SUM, comm->comm, ref->stream));
  CK(cudaStreamSynchronize(ref->stream));
  ref->covDirty = true;
  return T1K_OK;
}

int t1k_reads_load(const char *path1, const char *path2, T1KReads *out) {
  if (!path1 || !out) return fail(T1K_ERR_ARG, "t1k_reads_load: bad argument");
  memset(out, 0, sizeof(*out));
  LoadedReads R;
  std::string err;
  if (load_reads(path1, path2, T1K_MAX_READ_LEN, R, err)) return fail(T1K_ERR_ARG, "t1k_reads_load: " + err);
  const uint32_t n = (uint32_t)(R.off[0].size() - 1), stride = R.maxLen + 1;
  const int mates = path2 ? 2 : 1;
  int nDev = 0;
  const bool pin = cudaGetDeviceCount(&nDev) == cudaSuccess && nDev > 0;
  if (!pin) cudaGetLastError();
  char *buf[2] = {nullptr, nullptr};
  const size_t bytes = std::max<size_t>((size_t)n * stride, 16);
  for (int m = 0; m < mates; ++m) {
    if (pin) { if (cudaMallocHost((void **)&buf[m], bytes) != cudaSuccess) { cudaGetLastError(); buf[m] = nullptr; } }
    else buf[m] = (char *)malloc(bytes);
    if (!buf[m]) { for (int k = 0; k < m; ++k) { if (pin) cudaFreeHost(buf[k]); else free(buf[k]); } return fail(T1K_ERR_ARG, "t1k_reads_load: out of host memory"); }
  }
  const int T = n < 4096 ? 1 : (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  run_threads(T, [&](int t) {
    for (int m = 0; m < mates; ++m)
      for (size_t i = (size_t)n * t / T; i < (size_t)n * (t + 1) / T; ++i) {
        const size_t len = (size_t)(R.off[m][i + 1] - R.off[m][i]);
        char *dst = buf[m] + i * stride;
        memcpy(dst, R.bases[m].data() + R.off[m][i], len);
        memset(dst + len, 0, stride - len);
      }
  });
  out->reads1 = buf[0]; out->reads2 = buf[1]; out->stride = stride; out->n_frag = n; out->max_len = R.maxLen; out->pinned = pin ? 1 : 0;
  return T1K_OK;
}

void t1k_reads_free(T1KReads *r) {
  if (!r) return;
  if (r->pinned) { if (r->reads1) cudaFreeHost(r->reads1); if (r->reads2) cudaFreeHost(r->reads2); }
  else { free(r->reads1); free(r->reads2); }
  r->reads1 = r->reads2 = nullptr; r->n_frag = 0;
}

int t1k_groups_create(T1KGroups **out) {
  if (!out) return fail(T1K_ERR_ARG, "t1k_groups_create: bad argument");
  *out = new T1KGroups;
  return T1K_OK;
}

void t1k_groups_destroy(T1KGroups *g) { delete g; }

int t1k_groups_add_fragments(T1KGroups *g, const uint64_t *row_ptr, const T1KReadAssignment *entries, uint32_t n_frag) {
  if (!g || (n_frag > 0 && (!row_ptr || !entries))) return fail(T1K_ERR_ARG, "t1k_groups_add_fragments: bad argument");
  std::vector<HostEntry> row;
  for (uint32_t f = 0; f < n_frag; ++f) {
    const uint64_t b = row_ptr[f], e = row_ptr[f + 1];
    if (e <= b) continue;
    row.resize(e - b);
    memcpy(row.data(), entries + b, (e - b) * sizeof(HostEntry));
    // Genotyper.hpp:851: the fragment's assignments are sorted by allele before they are compared
    std::stable_sort(row.begin(), row.end(), [](const HostEntry &x, const HostEntry &y) { return x.alleleIdx < y.alleleIdx; });
    g->G.add(row.data(), (uint32_t)(e - b));
  }
  return T1K_OK;
}

int t1k_groups_serialize(const T1KGroups *g, void **blob, uint64_t *bytes) {
  if (!g || !blob || !bytes) return fail(T1K_ERR_ARG, "t1k_groups_serialize: bad argument");
  std::vector<uint8_t> b;
  serialize_groups(g->G, b);
  void *p = malloc(b.size());
  if (!p) return fail(T1K_ERR_ARG, "out of host memory");
  memcpy(p, b.data(), b.size());
  *blob = p; *bytes = b.size();
  return T1K_OK;
}

int t1k_groups_merge(T1KGroups *g, const void *blob, uint64_t bytes) {
  if (!g || !blob) return fail(T1K_ERR_ARG, "t1k_groups_merge: bad argument");
  if (!merge_groups(g->G, (const uint8_t *)blob, bytes)) return fail(T1K_ERR_ARG, "t1k_groups_merge: malformed group table");
  return T1K_OK;
}

int t1k_groups_fetch(const T1KGroups *g, int32_t *n_groups, uint64_t *n_entries, uint64_t *assigned_fragments, int64_t *ptr,
                     T1KReadAssignment *entries) {
  if (!g) return fail(T1K_ERR_ARG, "t1k_groups_fetch: bad argument");
  if (n_groups) *n_groups = g->G.size();
  if (n_entries) *n_entries = g->G.ent.size();
  if (assigned_fragments) *assigned_fragments = (uint64_t)g->G.assignedFragments;
  if (ptr) memcpy(ptr, g->G.ptr.data(), g->G.ptr.size() * 8);
  if (entries && !g->G.ent.empty()) memcpy(entries, g->G.ent.data(), g->G.ent.size() * sizeof(HostEntry));
  return T1K_OK;
}

int t1k_groups_ec_filter(const T1KGroups *g, int32_t n_alleles, const int32_t *allele_len, const double *ec_abundance, const int32_t *ec_allele_ptr,
                         const int32_t *ec_alleles, int32_t n_ec, uint8_t *allele_kept, int32_t *allele_span) {
  if (!g || n_alleles < 0 || n_ec < 0 || !allele_len || !ec_abundance || !ec_allele_ptr || !ec_alleles || !allele_kept)
    return fail(T1K_ERR_ARG, "t1k_groups_ec_filter: bad argument");
  for (size_t k = 0; k < g->G.ent.size(); ++k)
    if (g->G.ent[k].alleleIdx < 0 || g->G.ent[k].alleleIdx >= n_alleles) return fail(T1K_ERR_ARG, "t1k_groups_ec_filter: allele index out of range");
  for (int32_t k = 0; k < ec_allele_ptr[n_ec]; ++k)
    if (ec_alleles[k] < 0 || ec_alleles[k] >= n_alleles) return fail(T1K_ERR_ARG, "t1k_groups_ec_filter: class member out of range");
  std::vector<int32_t> spans;
  allele_spans(g->G, n_alleles, std::max(1u, std::min(16u, std::thread::hardware_concurrency())), spans);
  ec_likelihood_filter(ec_allele_ptr, ec_alleles, n_ec, n_alleles, allele_len, ec_abundance, spans.data(), allele_kept);
  if (allele_span) memcpy(allele_span, spans.data(), spans.size() * 4);
  return T1K_OK;
}

int t1k_em_partition(const int64_t *row_ptr, int32_t n_groups, int32_t world, int32_t *bounds) {
  if (!row_ptr || !bounds || n_groups < 0 || world < 1) return fail(T1K_ERR_ARG, "t1k_em_partition: bad argument");
  partition_rows(row_ptr, n_groups, world, bounds);
  return T1K_OK;
}

}  // extern "C"

namespace {

// all-gather of one host blob per rank through device staging (padded to the largest); the result lands in pinned
// memory `recv`: blob r = recv + r * stride, sizes[r] bytes
int allgather_blobs(T1KRef *ref, T1KComm *comm, const uint8_t *mine, uint64_t myBytes, PinnedMem &recv, uint64_t &stride,
                    std::vector<uint64_t> &sizes) {
  cudaStream_t st = ref->stream;
  const int W = comm->world;
  DevMem dSizes, dBuf;
  CK(dSizes.alloc((size_t)W * 8));
  CK(cudaMemcpyAsync(dSizes.as<uint64_t>() + comm->rank, &myBytes, 8, cudaMemcpyHostToDevice, st));
  NK(nccl().AllGather(dSizes.as<uint64_t>() + comm->rank, dSizes.p, 1, NCCL_UINT64, comm->comm, st));
  sizes.assign(W, 0);
  CK(cudaMemcpyAsync(sizes.data(), dSizes.p, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  uint64_t mx = 16;
  for (int r = 0; r < W; ++r) mx = std::max(mx, sizes[r]);
  mx = (mx + 15) & ~15ull;
  stride = mx;
  CK(dBuf.alloc((size_t)W * mx));
  uint8_t *slot = dBuf.as<uint8_t>() + (size_t)comm->rank * mx;
  if (myBytes) CK(cudaMemcpyAsync(slot, mine, myBytes, cudaMemcpyHostToDevice, st));
  NK(nccl().AllGather(slot, dBuf.p, mx, NCCL_UINT8, comm->comm, st));
  CK(recv.grow((size_t)W * mx, 0));
  for (int r = 0; r < W; ++r)
    if (sizes[r]) CK(cudaMemcpyAsync(recv.as<uint8_t>() + (size_t)r * mx, dBuf.as<uint8_t>() + (size_t)r * mx, sizes[r], cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return T1K_OK;
}

// The same all-gather for COMPACT partitions that the device tail consumes where they land: the blobs stay in `dBuf`
// (blob r at r * stride); only their heads ([nG, nE], ptr, first, count) are copied to pinned memory `recv` (head r at
// headOff[r], headBytes[r] bytes).
int allgather_blobs_dev(T1KRef *ref, T1KComm *comm, const uint8_t *mine, uint64_t myBytes, DevMem &dBuf, PinnedMem &recv, uint64_t &stride,
                        std::vector<uint64_t> &sizes, std::vector<uint64_t> &headOff, std::vector<uint64_t> &headBytes) {
  cudaStream_t st = ref->stream;
  const int W = comm->world;
  DevMem dSizes;
  CK(dSizes.alloc((size_t)W * 8));
  CK(cudaMemcpyAsync(dSizes.as<uint64_t>() + comm->rank, &myBytes, 8, cudaMemcpyHostToDevice, st));
  NK(nccl().AllGather(dSizes.as<uint64_t>() + comm->rank, dSizes.p, 1, NCCL_UINT64, comm->comm, st));
  sizes.assign(W, 0);
  CK(cudaMemcpyAsync(sizes.data(), dSizes.p, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  uint64_t mx = 16;
  for (int r = 0; r < W; ++r) mx = std::max(mx, sizes[r]);
  mx = (mx + 15) & ~15ull;
  stride = mx;
  CK(dBuf.alloc((size_t)W * mx));
  uint8_t *slot = dBuf.as<uint8_t>() + (size_t)comm->rank * mx;
  if (myBytes) CK(cudaMemcpyAsync(slot, mine, myBytes, cudaMemcpyHostToDevice, st));
  NK(nccl().AllGather(slot, dBuf.p, mx, NCCL_UINT8, comm->comm, st));
  std::vector<uint64_t> hdr((size_t)W * 2, 0);
  for (int r = 0; r < W; ++r)
    if (sizes[r] >= 16) CK(cudaMemcpyAsync(&hdr[(size_t)r * 2], dBuf.as<uint8_t>() + (size_t)r * mx, 16, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  headOff.assign(W, 0); headBytes.assign(W, 0);
  uint64_t tot = 0;
  for (int r = 0; r < W; ++r) {
    const uint64_t want = 16 + (hdr[(size_t)r * 2] + 1) * 8 + hdr[(size_t)r * 2] * 16;
    headBytes[r] = std::min<uint64_t>(want, sizes[r]);
    headOff[r] = tot; tot += (headBytes[r] + 15) & ~15ull;
  }
  CK(recv.grow(std::max<size_t>(tot, 16), 0));
  for (int r = 0; r < W; ++r)
    if (headBytes[r]) CK(cudaMemcpyAsync(recv.as<uint8_t>() + headOff[r], dBuf.as<uint8_t>() + (size_t)r * mx, headBytes[r], cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return T1K_OK;
}

// every rank contributes `n` words; all[r * n + i] = word i of rank r
int allgather_u64(T1KRef *ref, T1KComm *comm, const uint64_t *mine, int n, std::vector<uint64_t> &all) {
  cudaStream_t st = ref->stream;
  const int W = comm->world;
  DevMem d;
  CK(d.alloc((size_t)W * n * 8));
  CK(cudaMemcpyAsync(d.as<uint64_t>() + (size_t)comm->rank * n, mine, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  NK(nccl().AllGather(d.as<uint64_t>() + (size_t)comm->rank * n, d.p, (size_t)n, NCCL_UINT64, comm->comm, st));
  all.assign((size_t)W * n, 0);
  CK(cudaMemcpyAsync(all.data(), d.p, (size_t)W * n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return T1K_OK;
}

// all-to-all of host blobs through device staging over NVLink (ncclSend / ncclRecv in one group): `send` holds this rank's
// blobs for rank 0, 1, ... back to back (sendBytes[r] each, multiples of 16); bytesFrom[s] = what rank s sends to this
// rank (from the size exchange); the incoming blobs land back to back in pinned memory `recv`.
int alltoall_blobs(T1KRef *ref, T1KComm *comm, const uint8_t *send, const std::vector<size_t> &sendBytes, const std::vector<uint64_t> &bytesFrom,
                   PinnedMem &recv, std::vector<size_t> &recvOff) {
  cudaStream_t st = ref->stream;
  const int W = comm->world;
  size_t totSend = 0, totRecv = 0;
  std::vector<size_t> sendOff((size_t)W + 1, 0);
  recvOff.assign((size_t)W + 1, 0);
  for (int r = 0; r < W; ++r) { sendOff[r + 1] = sendOff[r] + sendBytes[r]; recvOff[r + 1] = recvOff[r] + (size_t)bytesFrom[r]; }
  totSend = sendOff[W]; totRecv = recvOff[W];
  DevMem dSend, dRecv;
  CK(dSend.alloc(std::max<size_t>(totSend, 16))); CK(dRecv.alloc(std::max<size_t>(totRecv, 16)));
  if (totSend) CK(cudaMemcpyAsync(dSend.p, send, totSend, cudaMemcpyHostToDevice, st));
  NK(nccl().GroupStart());
  for (int r = 0; r < W; ++r) {
    if (sendBytes[r]) NK(nccl().Send(dSend.as<uint8_t>() + sendOff[r], sendBytes[r], NCCL_UINT8, r, comm->comm, st));
    if (bytesFrom[r]) NK(nccl().Recv(dRecv.as<uint8_t>() + recvOff[r], (size_t)bytesFrom[r], NCCL_UINT8, r, comm->comm, st));
  }
  NK(nccl().GroupEnd());
  CK(recv.grow(std::max<size_t>(totRecv, 16), 0));
  if (totRecv) CK(cudaMemcpyAsync(recv.p, dRecv.p, totRecv, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return T1K_OK;
}

}  // namespace

namespace {

// ---- the global tail on the device (t1k_tail.cuh)
struct TailDevice {
  DevMem gPtr, gAllele, bits, fp, rowHash, listLen, alleleEc, isRep, ecRep, rowLen, rowPtr, col, colLen, colBeg, rowIdx, pairs, differ;
  TailParams P;
  std::vector<int64_t> hRowPtr;      // the EM's CSR row pointer (host copy: row partition, nnz)
};

// Equivalence classes of the alleles from the bit matrix.  *ok = false: a (fingerprint, hash, length) bucket held alleles with
// different lists (a 64-bit hash collision) — nothing is lost, the caller takes the host path.
int device_tail_classes(T1KRef *ref, const GroupsView &G, int32_t nA, int threads, TailDevice &T, EquivalenceClasses &EC, bool *ok) {
  *ok = false;
  cudaStream_t st = ref->stream;
  const int32_t n = G.n;
  const int64_t nE = G.entries();
  // allele ids of the entries: the compact form has them as they are; a full table is read once
  PhaseTimer pt;
  const int32_t *hAllele = G.allele;
  if (!hAllele && !G.dAllele) {
    CK(ref->pinIds.ensure((size_t)std::max<int64_t>(nE, 1) * 4));      // (pinned: the upload is a DMA at link speed)
    int32_t *ids = ref->pinIds.as<int32_t>();
    if (threads < 1 || (size_t)nE < par_min_entries()) threads = 1;
    const HostEntry *ent = G.ent;
    run_threads(threads, [&](int t) { for (int64_t k = nE * t / threads; k < nE * (t + 1) / threads; ++k) ids[k] = ent[k].alleleIdx; });
    hAllele = ids;
  }
  pt.lap("  tail: allele ids");
  const int32_t W = (n + 63) / 64;
  CK(T.gPtr.alloc(((size_t)n + 1) * 8));
  if (!G.dAllele) CK(T.gAllele.alloc((size_t)std::max<int64_t>(nE, 1) * 4));
  CK(T.bits.alloc((size_t)nA * std::max(W, 1) * 8));
  CK(T.fp.alloc((size_t)nA * 4)); CK(T.rowHash.alloc((size_t)nA * 8)); CK(T.listLen.alloc((size_t)nA * 4)); CK(T.differ.alloc(4));
  CK(cudaMemcpyAsync(T.gPtr.p, G.ptr, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
  if (nE && !G.dAllele) CK(cudaMemcpyAsync(T.gAllele.p, hAllele, (size_t)nE * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(T.bits.p, 0, (size_t)nA * std::max(W, 1) * 8, st));
  CK(cudaMemsetAsync(T.differ.p, 0, 4, st));
  TailParams &P = T.P;
  memset(&P, 0, sizeof(P));
  P.nGroups = n; P.nAlleles = nA; P.gPtr = T.gPtr.as<int64_t>(); P.gAllele = G.dAllele ? G.dAllele : T.gAllele.as<int32_t>();
  P.bits = T.bits.as<u64>(); P.wordsPerRow = std::max(W, 1);
  P.fp = T.fp.as<int32_t>(); P.rowHash = T.rowHash.as<u64>(); P.listLen = T.listLen.as<int32_t>();
  k_tail_bits<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, st>>>(P);
  k_tail_fp<<<(nA + 127) / 128, 128, 0, st>>>(P);
  CK(cudaGetLastError());
  std::vector<int32_t> fp(nA), len(nA);
  std::vector<u64> hs(nA);
  CK(cudaMemcpyAsync(fp.data(), T.fp.p, (size_t)nA * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(len.data(), T.listLen.p, (size_t)nA * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hs.data(), T.rowHash.p, (size_t)nA * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  pt.lap("  tail: upload + bit matrix + fingerprints");
  // the reference's order (Genotyper.hpp:1094-1101): fingerprint descending, allele ascending; alleles without reads last
  std::vector<int32_t> order(nA);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return fp[x] != fp[y] ? fp[y] < fp[x] : x < y; });
  // candidates for one class: same fingerprint, row hash and list length.  Inside a fingerprint run the buckets keep the order
  // of their first members, which is the order the reference creates the classes in.
  EC.alleleEc.assign(nA, -1);
  std::vector<std::vector<int32_t> > ecs;
  std::vector<int32_t> pairs;
  for (int32_t i = 0; i < nA;) {
    if (fp[order[i]] == -1) break;
    int32_t j = i;
    while (j < nA && fp[order[j]] == fp[order[i]]) ++j;
    const size_t firstClass = ecs.size();
    for (int32_t k = i; k < j; ++k) {
      const int32_t a = order[k];
      size_t c = firstClass;
      for (; c < ecs.size(); ++c) { const int32_t r = ecs[c][0]; if (hs[r] == hs[a] && len[r] == len[a]) break; }
      if (c == ecs.size()) ecs.push_back(std::vector<int32_t>(1, a));
      else { pairs.push_back(ecs[c][0]); pairs.push_back(a); ecs[c].push_back(a); }
      EC.alleleEc[a] = (int32_t)c;
    }
    i = j;
  }
  pt.lap("  tail: class buckets (host)");
  if (!pairs.empty()) {
    const int nPairs = (int)(pairs.size() / 2);
    CK(T.pairs.alloc(pairs.size() * 4));
    CK(cudaMemcpyAsync(T.pairs.p, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice, st));
    k_tail_cmp<<<(unsigned)(((size_t)nPairs * 32 + 255) / 256), 256, 0, st>>>(P, T.pairs.as<int32_t>(), nPairs, T.differ.as<int>());
    CK(cudaGetLastError());
    int differ = 0;
    CK(cudaMemcpyAsync(&differ, T.differ.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (differ) { EC.alleleEc.clear(); return T1K_OK; }
  }
  EC.ecPtr.assign(1, 0); EC.ecAlleles.clear();
  for (size_t e = 0; e < ecs.size(); ++e) {
    EC.ecAlleles.insert(EC.ecAlleles.end(), ecs[e].begin(), ecs[e].end());
    EC.ecPtr.push_back((int32_t)EC.ecAlleles.size());
  }
  *ok = true;
  return T1K_OK;
}

// The EM's matrix from the bit matrix and the classes: CSR of all rows (first-appearance order inside a row), CSC of the rows
// [g0, g1) this rank's E-step covers (the partition needs the row pointer, so the CSR comes first).
int device_tail_matrix(T1KRef *ref, TailDevice &T, const EquivalenceClasses &EC, int32_t nA, T1KComm *comm, EmDevice &out) {
  cudaStream_t st = ref->stream;
  TailParams &P = T.P;
  const int32_t n = P.nGroups, E = EC.size();
  std::vector<u8> isRep(nA, 0);
  std::vector<int32_t> ecRep(std::max(E, 1));
  for (int32_t e = 0; e < E; ++e) { ecRep[e] = EC.ecAlleles[EC.ecPtr[e]]; isRep[ecRep[e]] = 1; }
  CK(T.alleleEc.alloc((size_t)nA * 4)); CK(T.isRep.alloc(nA)); CK(T.ecRep.alloc((size_t)std::max(E, 1) * 4));
  CK(T.rowLen.alloc((size_t)std::max(n, 1) * 8)); CK(T.rowPtr.alloc(((size_t)n + 1) * 8));
  CK(T.colLen.alloc((size_t)std::max(E, 1) * 8)); CK(T.colBeg.alloc(((size_t)E + 1) * 8));
  CK(cudaMemcpyAsync(T.alleleEc.p, EC.alleleEc.data(), (size_t)nA * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(T.isRep.p, isRep.data(), nA, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(T.ecRep.p, ecRep.data(), (size_t)std::max(E, 1) * 4, cudaMemcpyHostToDevice, st));
  P.nEc = E; P.alleleEc = T.alleleEc.as<int32_t>(); P.isRep = T.isRep.as<u8>(); P.ecRep = T.ecRep.as<int32_t>();
  P.rowLen = T.rowLen.as<int64_t>(); P.rowPtr = T.rowPtr.as<int64_t>(); P.colLen = T.colLen.as<int64_t>(); P.colBeg = T.colBeg.as<int64_t>();
  const unsigned gGroups = (unsigned)(((size_t)n * 32 + 255) / 256), gEc = (unsigned)(((size_t)E * 32 + 255) / 256);
  k_tail_rowlen<<<gGroups, 256, 0, st>>>(P);
  k_scan_i64<<<1, 1024, 0, st>>>(P.rowLen, n, P.rowPtr);
  CK(cudaGetLastError());
  T.hRowPtr.resize((size_t)n + 1);
  CK(cudaMemcpyAsync(T.hRowPtr.data(), T.rowPtr.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int64_t nnz = T.hRowPtr[n];
  CK(T.col.alloc((size_t)std::max<int64_t>(nnz, 1) * 4));
  P.col = T.col.as<int32_t>();
  k_tail_rowfill<<<gGroups, 256, 0, st>>>(P);
  int g0 = 0, g1 = n;
  if (comm) {
    std::vector<int32_t> bounds((size_t)comm->world + 1);
    partition_rows(T.hRowPtr.data(), n, comm->world, bounds.data());
    g0 = bounds[comm->rank]; g1 = bounds[comm->rank + 1];
  }
  P.g0 = g0; P.g1 = g1;
  k_tail_collen<<<gEc, 256, 0, st>>>(P);
  k_scan_i64<<<1, 1024, 0, st>>>(P.colLen, E, P.colBeg);
  CK(cudaGetLastError());
  int64_t nnzL = 0;
  CK(cudaMemcpyAsync(&nnzL, T.colBeg.as<int64_t>() + E, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(T.rowIdx.alloc((size_t)std::max<int64_t>(nnzL, 1) * 4));
  P.rowIdx = T.rowIdx.as<int32_t>();
  k_tail_colfill<<<gEc, 256, 0, st>>>(P);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  out.rowPtr = T.rowPtr.as<int64_t>(); out.col = T.col.as<int32_t>();
  out.colBeg = T.colBeg.as<int64_t>(); out.colEnd = T.colBeg.as<int64_t>() + 1; out.rowIdx = T.rowIdx.as<int32_t>();
  out.g0 = g0; out.g1 = g1;
  return T1K_OK;
}

}  // namespace

extern "C" {

// Genotyper.cpp:450-646 for one sample.
int t1k_genotype(T1KRef *ref, const char *reads1, const char *reads2, uint32_t stride, uint32_t n_frag,
                 const T1KGenotypeParams *prm, T1KGenotypeResult *res) {
  if (!ref || !reads1 || !prm || !res || stride == 0 || !prm->effective_len) return fail(T1K_ERR_ARG, "t1k_genotype: bad argument");
  CK(cudaSetDevice(ref->device));
  const int32_t nA = ref->nAlleles;
  // ---- fragments in chunks through a three-stage pipeline (host prep | device align + pair | host coalesce):
  //   prep      unique read-ends of the chunk (Genotyper.cpp:450-454 sorts; only the grouping matters) + batch layout
  //   device    t1k_assign_batch + pairing kernels, rows copied into pinned memory
  //   coalesce  Genotyper::CoalesceReadAssignments, serial and in chunk order
  // While the device works on chunk c, helper threads prepare chunk c+1 and coalesce chunk c-1.
  u32 chunk = 1u << 18;
  if (const char *env = getenv("T1K_CHUNK_FRAGMENTS")) chunk = (u32)std::max(1l, atol(env));
  typedef ReadEndChunk Prep;
  Prep prep[2];
  int prepThreads = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() / ((prm->comm && prm->comm->world > 1) ? prm->comm->world : 1) / 2));
  if (const char *env = getenv("T1K_PREP_THREADS")) prepThreads = std::max(1, atoi(env));
  auto do_prep = [&, prepThreads](Prep &C, u32 f0, u32 m) {
    const double t = now_ms();
    unique_read_ends(reads1, reads2, stride, f0, m, T1K_MAX_READ_LEN, prepThreads, C);
    C.ms = now_ms() - t;
  };
  // Device-side ingest (t1k_ingest.cuh): the prepare stage shrinks to one H2D copy of the chunk's raw reads, packing
  // and de-duplication run on the device in front of the AssignRead kernels.  Reads up to 255 bases (one plane geometry
  // whatever the reads' actual lengths); longer strides and T1K_HOST_DEDUP=1 take the host de-duplication above.
  const bool devIngest = stride <= 255 && getenv("T1K_HOST_DEDUP") == nullptr;
  struct DevChunk {
    DevMem raw1, raw2;
    u32 f0 = 0, m = 0; double ms = 0; int rc = 0; std::string err;
  } dchunk[2];
  auto do_prep_dev = [&](DevChunk &C, u32 f0, u32 m) {
    const double t = now_ms();
    C.f0 = f0; C.m = m; C.rc = 0;
    cudaError_t e = cudaSetDevice(ref->device);
    if (e == cudaSuccess) e = C.raw1.alloc((size_t)m * stride);
    if (e == cudaSuccess) e = cudaMemcpyAsync(C.raw1.p, reads1 + (size_t)f0 * stride, (size_t)m * stride, cudaMemcpyHostToDevice, ref->prepStream);
    if (e == cudaSuccess && reads2) e = C.raw2.alloc((size_t)m * stride);
    if (e == cudaSuccess && reads2) e = cudaMemcpyAsync(C.raw2.p, reads2 + (size_t)f0 * stride, (size_t)m * stride, cudaMemcpyHostToDevice, ref->prepStream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ref->prepStream);
    if (e != cudaSuccess) { C.rc = T1K_ERR_CUDA; C.err = std::string("H2D of the chunk's reads: ") + cudaGetErrorString(e); }
    C.ms = now_ms() - t;
  };
  // packing + de-duplication of one chunk on ref->stream -> compact batch (DevReads) + per-fragment (end1, end2, hasN) on the device
  struct DevBatch { DevMem planesAll, lenAll, endHasN, hash, table, repOf, cnt, uid, blockSum, planes, len16, w, e1, e2, fragHasN, nUnique; u32 nU = 0; int maxLen = 0; float ms = 0; };
  auto ingest_chunk = [&](DevChunk &C, DevBatch &B) -> int {
    cudaStream_t st = ref->stream;
    const int mates = reads2 ? 2 : 1;
    const u32 nEnds = C.m * (u32)mates;
    if (int rc = setup_assign_launch(ref, KMER)) return rc;          // (geometry for reads up to 255 bases; refined below)
    const int RW = read_words(255);
    u32 tab = 1024; while (tab < 2 * nEnds) tab <<= 1;
    CK(B.planesAll.alloc((size_t)nEnds * 4 * RW * 8)); CK(B.lenAll.alloc((size_t)nEnds * 2)); CK(B.endHasN.alloc(nEnds)); CK(B.hash.alloc((size_t)nEnds * 8));
    CK(B.table.alloc((size_t)tab * 4)); CK(B.repOf.alloc((size_t)nEnds * 4)); CK(B.cnt.alloc((size_t)nEnds * 4)); CK(B.uid.alloc((size_t)nEnds * 4));
    const u32 nBlocks = (nEnds + 1023) / 1024;
    CK(B.blockSum.alloc((size_t)std::max(1u, nBlocks) * 4));
    CK(B.planes.alloc((size_t)nEnds * 4 * RW * 8)); CK(B.len16.alloc((size_t)nEnds * 2)); CK(B.w.alloc((size_t)nEnds * 4));
    CK(B.e1.alloc((size_t)C.m * 4)); CK(B.e2.alloc((size_t)C.m * 4)); CK(B.fragHasN.alloc(C.m)); CK(B.nUnique.alloc(8));
    CK(cudaMemsetAsync(B.table.p, 0xff, (size_t)tab * 4, st));
    CK(cudaMemsetAsync(B.nUnique.p, 0, 8, st));
    CK(cudaMemsetAsync(ref->errFlag.p, 0, sizeof(int), st));
    CK(cudaMemsetAsync(ref->stats.p, 0, 8 * sizeof(unsigned long long), st));
    IngestParams P;
    P.raw1 = C.raw1.as<char>(); P.raw2 = reads2 ? C.raw2.as<char>() : nullptr; P.stride = stride; P.m = C.m; P.nEnds = nEnds; P.mates = mates; P.RW = RW;
    P.planesAll = B.planesAll.as<u64>(); P.lenAll = B.lenAll.as<u16>(); P.endHasN = B.endHasN.as<u8>(); P.hash = B.hash.as<u64>();
    P.table = B.table.as<u32>(); P.tabMask = tab - 1; P.repOf = B.repOf.as<u32>(); P.cnt = B.cnt.as<u32>(); P.uid = B.uid.as<u32>(); P.blockSum = B.blockSum.as<u32>();
    P.planes = B.planes.as<u64>(); P.len16 = B.len16.as<u16>(); P.w = B.w.as<int32_t>();
    P.e1 = B.e1.as<u32>(); P.e2 = B.e2.as<u32>(); P.fragHasN = B.fragHasN.as<u8>(); P.nUnique = B.nUnique.as<u32>(); P.err = ref->errFlag.as<int>();
    cudaEvent_t ev0, ev1;
    CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
    CK(cudaEventRecord(ev0, st));
    if (nEnds) {
      const unsigned g = (nEnds + 255) / 256;
      k_ingest_pack<<<g, 256, 0, st>>>(P);
      k_dedup_insert<<<g, 256, 0, st>>>(P);
      k_dedup_resolve<<<g, 256, 0, st>>>(P);
      k_scan_block<<<nBlocks, 1024, 0, st>>>(P);
      k_scan_sums<<<1, 1024, 0, st>>>(P, nBlocks);
      k_dedup_emit<<<g, 256, 0, st>>>(P);
      k_dedup_map<<<(C.m + 255) / 256, 256, 0, st>>>(P);
      CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ev1, st));
    u32 h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, B.nUnique.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&B.ms, ev0, ev1));
    B.nU = h[0]; B.maxLen = (int)h[1];
    return setup_assign_launch(ref, std::max<int>(KMER, B.maxLen));
  };
  CK(cudaMemsetAsync(ref->covDiff.p, 0, ref->covEntries * 4, ref->stream));
  CK(cudaMemsetAsync(ref->covPoint.p, 0, ref->covEntries * 4, ref->stream));
  ref->covDirty = true;
  ReadGroups groups;
  int hostThreads = (int)std::thread::hardware_concurrency() / ((prm->comm && prm->comm->world > 1) ? prm->comm->world : 1);
  if (const char *env = getenv("T1K_HOST_THREADS")) hostThreads = atoi(env);
  GroupShards shards(std::max(1, std::min(8, hostThreads)));
  res->n_unique_ends = 0; res->n_overlaps = 0; res->n_assignments = 0;
  res->ms_dedup = res->ms_align = res->ms_pair = res->ms_coalesce = res->ms_em = 0;
  res->ms_align_kernel = res->ms_pair_kernel = res->ms_em_kernel = 0;
  res->n_postings = res->n_candidates = 0; res->n_launches = 0;
  res->ms_prep_wait = res->ms_exchange = 0; res->n_pair_records = 0; res->em_nnz = 0; res->em_updates = 0;
  PairHost pairOut[2];
  pairOut[0].pin = &ref->pinEntries[0]; pairOut[1].pin = &ref->pinEntries[1];
  double msCoalesce = 0;
  int coalRc = 0;
  auto do_coalesce = [&](PairHost &H, u32 f0, u32 m) {
    if (H.finish() != cudaSuccess) { coalRc = 1; return; }       // the rows of this chunk arrive on the copy stream
    const double t = now_ms();
    shards.add_chunk(H.entries(), H.rowOff.data(), H.rowCnt.data(), H.rowHash.data(), m, (int64_t)f0);
    if (res->fragment_assigned) memcpy(res->fragment_assigned + f0, H.assigned.data(), m);
    msCoalesce += now_ms() - t;
  };
  std::thread prepThread, coalThread;
  struct Joiner { std::thread &a, &b; ~Joiner() { if (a.joinable()) a.join(); if (b.joinable()) b.join(); } } joiner{prepThread, coalThread};
  // chunk boundaries: the first chunk is small, because its preparation is the one stage nothing can overlap with
  std::vector<std::pair<u32, u32>> chunks;      // (first fragment, count)
  {
    u32 f0 = 0;
    u32 firstChunk = chunk;        // (a smaller first chunk starts the device earlier but de-duplicates less: T1K_FIRST_CHUNK)
    if (const char *env = getenv("T1K_FIRST_CHUNK")) firstChunk = (u32)std::max(1l, atol(env));
    while (f0 < n_frag) {
      const u32 m = std::min(chunks.empty() ? firstChunk : chunk, n_frag - f0);
      chunks.push_back(std::make_pair(f0, m));
      f0 += m;
    }
  }
  const u32 nChunks = (u32)chunks.size();
  double msPrepWait = 0;
  uint64_t nPairRecords = 0;
  // the rank-local stage (no collective inside): returns this rank's status instead of leaving early, so that in a
  // read-sharded run every rank reaches the status exchange below
  auto local_stage = [&]() -> int {
  {
    const double t = now_ms();
    if (nChunks) { if (devIngest) do_prep_dev(dchunk[0], chunks[0].first, chunks[0].second); else do_prep(prep[0], chunks[0].first, chunks[0].second); }
    msPrepWait += now_ms() - t;
  }
  for (u32 c = 0; c < nChunks; ++c) {
    Prep &C = prep[c & 1];
    DevChunk &DC = dchunk[c & 1];
    DevBatch DB;
    if (devIngest) { C.f0 = DC.f0; C.m = DC.m; C.ms = DC.ms; }
    res->ms_dedup += (float)C.ms;
    if (c + 1 < nChunks) {
      if (devIngest) prepThread = std::thread(do_prep_dev, std::ref(dchunk[(c + 1) & 1]), chunks[c + 1].first, chunks[c + 1].second);
      else prepThread = std::thread(do_prep, std::ref(prep[(c + 1) & 1]), chunks[c + 1].first, chunks[c + 1].second);
    }
    double ta = now_ms();
    T1KAssignment *a = nullptr;
    if (devIngest) {
      if (DC.rc) { g_err = DC.err; return DC.rc; }
      if (int rc = ingest_chunk(DC, DB)) return rc;
      res->ms_dedup += DB.ms; res->n_launches += 7;
      res->n_unique_ends += DB.nU;
      const DevReads D = {DB.planes.as<u64>(), DB.len16.as<u16>(), DB.w.as<int32_t>(), DB.nU};
      if (int rc = assign_core(ref, D, &a)) return rc;
    } else {
      if (C.tooLong) return fail(T1K_ERR_ARG, "read longer than T1K_MAX_READ_LEN");
      res->n_unique_ends += C.rep.size();
      if (int rc = t1k_assign_batch(ref, C.bases.data(), C.off.data(), C.len.data(), C.w.data(), (uint32_t)C.rep.size(), &a)) return rc;
    }
    struct AG { T1KAssignment *a; ~AG() { t1k_assignment_destroy(a); } } ag{a};
    res->ms_align += (float)(now_ms() - ta);
    res->n_overlaps += a->storeUsed;
    res->ms_align_kernel += a->msKernel; res->n_postings += a->stats[0]; res->n_candidates += a->stats[1];
    res->n_launches += 2 + a->launches;
    double tp = now_ms();
    PairHost &H = pairOut[c & 1];     // last read by the coalescing of chunk c-2, which has been joined
    if (devIngest) {
      if (int rc = pair_fragments(ref, a, DB.e1.as<u32>(), reads2 ? DB.e2.as<u32>() : nullptr, DB.fragHasN.as<u8>(), C.m, prm->max_assign, false, H, true,
                                  getenv("T1K_SYNC_ROWS") == nullptr, true)) return rc;
    } else if (int rc = pair_fragments(ref, a, C.e1.data(), reads2 ? C.e2.data() : nullptr, C.hasN.data(), C.m, prm->max_assign, false, H, true,
                                       getenv("T1K_SYNC_ROWS") == nullptr)) return rc;
    res->ms_pair += (float)(now_ms() - tp);
    res->ms_pair_kernel += H.msKernel; res->n_launches += H.launches;
    res->n_assignments += H.nEntries;
    nPairRecords += H.nPairRecords;
    if (coalThread.joinable()) coalThread.join();
    if (coalRc) return fail(T1K_ERR_CUDA, "D2H of a chunk's fragment rows failed");
    coalThread = std::thread(do_coalesce, std::ref(H), C.f0, C.m);
    { const double t = now_ms(); if (prepThread.joinable()) prepThread.join(); msPrepWait += now_ms() - t; }
  }
  if (coalThread.joinable()) coalThread.join();
  if (coalRc) return fail(T1K_ERR_CUDA, "D2H of a chunk's fragment rows failed");
  return T1K_OK;
  };
  int localRc = local_stage();
  std::string localErr = g_err;
  if (prepThread.joinable()) prepThread.join();
  if (coalThread.joinable()) coalThread.join();
  T1KComm *comm = (prm->comm && prm->comm->world > 1) ? prm->comm : nullptr;
  if (!comm && localRc) return localRc;
  if (!localRc && !comm) {          // (a read-sharded run sends the shards' groups straight to their owners)
    const double t = now_ms();
    shards.gather(groups);
    msCoalesce += now_ms() - t;
  }
  res->ms_coalesce = (float)msCoalesce;
  res->ms_prep_wait = (float)msPrepWait; res->n_pair_records = nPairRecords; res->ms_exchange = 0;
  // ---- read-sharded run: total coverage; the read-group tables merged with the work divided over the
  // ranks: the hash space of the allele sets is cut into world x T partitions, every rank sends each peer the groups of the
  // partitions that peer owns (all-to-all over NVLink), merges its own partitions from all ranks in rank order (float32
  // weights add as per-rank partial sums, as when the ranks' tables are merged one after the other), and the merged
  // partitions are all-gathered and interleaved by the fragment that created each group = the single-process order.
  uint64_t nAssignAll = res->n_assignments;
  CompactGroups compactGroups;
  bool useCompact = false, compactOnDevice = false;
  DevMem dBlobs, dGPtr, dGAllele;     // read-sharded run: merged partitions / assembled (ptr, allele ids) resident on the device
  std::vector<int32_t> spans;       // N4: covered range of every allele over the coalesced groups
  if (comm) {
    double tx = now_ms();
    PhaseTimer px;
    const int W = comm->world, T = shards.threads();
    // status + sizes: {status, assignments, fragments, bytes for rank 0 .. W-1}
    ShardPlan plan;
    int64_t assignedLocal = 0;
    for (int t = 0; t < shards.threads(); ++t) assignedLocal += shards.part[t].assignedFragments;
    if (!localRc) plan_partitions(shards, W, T, plan);
    else { plan.bytes.assign((size_t)W, 0); plan.groupsOf.assign((size_t)W, std::vector<std::pair<int32_t, int32_t> >()); plan.total = 0; }
    std::vector<uint64_t> mineW((size_t)W + 4, 0), allW;
    mineW[0] = (uint64_t)localRc; mineW[1] = nAssignAll; mineW[2] = (uint64_t)n_frag; mineW[3] = (uint64_t)assignedLocal;
    for (int r = 0; r < W; ++r) mineW[4 + r] = (uint64_t)plan.bytes[r];
    if (int rc = allgather_u64(ref, comm, mineW.data(), W + 4, allW)) return rc;
    for (int r = 0; r < W; ++r)
      if (allW[(size_t)r * (W + 4)] != 0) {
        if (localRc) { g_err = localErr; return localRc; }
        return fail(T1K_ERR_NCCL, "rank " + std::to_string(r) + " of the read-sharded run failed before the exchange");
      }
    px.lap("exchange: plan + status");
    if (int rc = t1k_coverage_allreduce(ref, comm)) return rc;
    px.lap("exchange: coverage all-reduce");
    std::vector<int64_t> fragBase((size_t)W, 0);
    std::vector<uint64_t> bytesFrom((size_t)W, 0);
    nAssignAll = 0;
    int64_t fb = 0, assignedAll = 0;
    for (int r = 0; r < W; ++r) {
      const uint64_t *w = &allW[(size_t)r * (W + 4)];
      nAssignAll += w[1]; fragBase[r] = fb; fb += (int64_t)w[2]; assignedAll += (int64_t)w[3];
      bytesFrom[r] = w[4 + comm->rank];
    }
    CK(ref->pinSend.grow(std::max<size_t>(plan.total, 16), 0));
    serialize_partitions(shards, plan, ref->pinSend.as<uint8_t>(), T);
    px.lap("exchange: serialize partitions");
    std::vector<size_t> recvOff;
    if (int rc = alltoall_blobs(ref, comm, ref->pinSend.as<uint8_t>(), plan.bytes, bytesFrom, ref->pinRecv, recvOff)) return rc;
    px.lap("exchange: all-to-all");
    std::vector<GroupBlobView> tables((size_t)W);
    int mergeRc = 0;
    for (int r = 0; r < W && !mergeRc; ++r)
      if (!tables[r].parse(ref->pinRecv.as<uint8_t>() + recvOff[r], bytesFrom[r])) mergeRc = 1;
    ReadGroups mine, merged;
    if (!mergeRc && !merge_tables_partition(tables, fragBase, comm->rank, W, T, mine)) mergeRc = 1;
    px.lap("exchange: merge my partitions");
    if (!mergeRc && (res->allele_kept || res->allele_span)) allele_spans(mine, nA, T, spans);     // N4: this rank's share of the covered ranges
    // second exchange: status + the merged partitions.  The global tail (equivalence classes, EM inputs) reads only the allele ids
    // and the count of every group, so the partitions travel in compact form (4 bytes per entry instead of 24) unless the caller
    // wants the whole table back (groups_out).
    const bool compact = prm->groups_out == nullptr;
    const size_t partBytes = mergeRc ? 0 : (compact ? compact_group_bytes(mine) : serialized_group_bytes(mine));
    CK(ref->pinSend.grow(std::max<size_t>(partBytes, 16), 0));         // (the first exchange's send buffer is no longer needed)
    if (!mergeRc) { if (compact) serialize_compact(mine, ref->pinSend.as<uint8_t>(), T); else serialize_groups(mine, ref->pinSend.as<uint8_t>()); }
    px.lap("exchange: serialize merged");
    const uint64_t st2 = (uint64_t)mergeRc;
    std::vector<uint64_t> allSt;
    if (int rc = allgather_u64(ref, comm, &st2, 1, allSt)) return rc;
    for (int r = 0; r < W; ++r) if (allSt[r]) return fail(T1K_ERR_NCCL, "malformed read-group table on rank " + std::to_string(r));
    uint64_t stride2 = 0;
    std::vector<uint64_t> sizes2;
    const bool devAssemble = compact && getenv("T1K_HOST_TAIL") == nullptr;
    if (devAssemble) {
      // the merged partitions stay in HBM: the host sees their heads, the allele runs are gathered into the global order on the device
      std::vector<uint64_t> headOff, headBytes, blobBase((size_t)W);
      if (int rc = allgather_blobs_dev(ref, comm, ref->pinSend.as<uint8_t>(), partBytes, dBlobs, ref->pinRecv2, stride2, sizes2, headOff, headBytes)) return rc;
      px.lap("exchange: all-gather merged");
      std::vector<const uint8_t *> heads((size_t)W);
      for (int r = 0; r < W; ++r) { heads[r] = ref->pinRecv2.as<uint8_t>() + headOff[r]; blobBase[r] = (uint64_t)r * stride2; }
      std::vector<int64_t> srcOff;
      if (!assemble_compact_heads(heads, headBytes, sizes2, blobBase, compactGroups, srcOff)) return fail(T1K_ERR_NCCL, "malformed merged partition from a peer");
      const int32_t nG = (int32_t)compactGroups.ptr.size() - 1;
      const int64_t nE = compactGroups.ptr[nG];
      DevMem dSrc;
      CK(dSrc.alloc((size_t)std::max(nG, 1) * 8)); CK(dGPtr.alloc(((size_t)nG + 1) * 8)); CK(dGAllele.alloc((size_t)std::max<int64_t>(nE, 1) * 4));
      CK(cudaMemcpyAsync(dSrc.p, srcOff.data(), (size_t)nG * 8, cudaMemcpyHostToDevice, ref->stream));
      CK(cudaMemcpyAsync(dGPtr.p, compactGroups.ptr.data(), ((size_t)nG + 1) * 8, cudaMemcpyHostToDevice, ref->stream));
      if (nG) k_tail_gather<<<(unsigned)(((size_t)nG * 32 + 255) / 256), 256, 0, ref->stream>>>(dBlobs.as<uint8_t>(), dSrc.as<int64_t>(), dGPtr.as<int64_t>(), nG, dGAllele.as<int32_t>());
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(ref->stream));
      dBlobs.release();
      useCompact = true; compactOnDevice = true;
      groups.assignedFragments = assignedAll;
    } else {
    if (int rc = allgather_blobs(ref, comm, ref->pinSend.as<uint8_t>(), partBytes, ref->pinRecv2, stride2, sizes2)) return rc;
    px.lap("exchange: all-gather merged");
    }
    if (devAssemble) {
    } else if (compact) {
      std::vector<const uint8_t *> blobs((size_t)W);
      for (int r = 0; r < W; ++r) blobs[r] = ref->pinRecv2.as<uint8_t>() + (size_t)r * stride2;
      if (!assemble_compact(blobs, sizes2, T, compactGroups)) return fail(T1K_ERR_NCCL, "malformed merged partition from a peer");
      useCompact = true;
      groups.assignedFragments = assignedAll;
    } else {
      std::vector<GroupBlobView> parts((size_t)W);
      for (int r = 0; r < W; ++r)
        if (!parts[r].parse(ref->pinRecv2.as<uint8_t>() + (size_t)r * stride2, sizes2[r])) return fail(T1K_ERR_NCCL, "malformed merged partition from a peer");
      if (!assemble_partitions(parts, T, merged)) return fail(T1K_ERR_NCCL, "malformed merged partition from a peer");
      merged.assignedFragments = assignedAll;
      groups.ptr.swap(merged.ptr); groups.ent.swap(merged.ent); groups.byHash.swap(merged.byHash);
      groups.first.swap(merged.first); groups.hashes.swap(merged.hashes);
      groups.assignedFragments = merged.assignedFragments;
    }
    px.lap("exchange: assemble");
    if (res->allele_kept || res->allele_span) {
      // min / max over the ranks as ONE max all-reduce of {-minStart, maxEnd}
      if (spans.size() != (size_t)2 * nA) return fail(T1K_ERR_NCCL, "covered ranges missing on this rank");
      for (int32_t a = 0; a < nA; ++a) spans[a] = spans[a] == INT32_MAX ? INT32_MIN : -spans[a];
      DevMem dSp;
      CK(dSp.alloc(spans.size() * 4));
      CK(cudaMemcpyAsync(dSp.p, spans.data(), spans.size() * 4, cudaMemcpyHostToDevice, ref->stream));
      NK(nccl().AllReduce(dSp.p, dSp.p, spans.size(), NCCL_INT32, NCCL_MAX, comm->comm, ref->stream));
      CK(cudaMemcpyAsync(spans.data(), dSp.p, spans.size() * 4, cudaMemcpyDeviceToHost, ref->stream));
      CK(cudaStreamSynchronize(ref->stream));
      for (int32_t a = 0; a < nA; ++a) spans[a] = spans[a] == INT32_MIN ? INT32_MAX : -spans[a];
    }
    res->ms_exchange = (float)(now_ms() - tx);
    res->ms_coalesce += res->ms_exchange;
  } else if (res->allele_kept || res->allele_span) {
    allele_spans(groups, nA, shards.threads(), spans);
  }
  res->assigned_fragments = (int32_t)groups.assignedFragments;
  // Genotyper::GetAverageReadAssignmentCnt (Genotyper.hpp:941-955) averages over the coalesced read groups
  GroupsView GV = useCompact ? compactGroups.view() : view_of(groups);
  if (compactOnDevice) { GV.allele = nullptr; GV.dAllele = dGAllele.as<int32_t>(); }
  res->avg_alleles_per_read = GV.n ? (double)GV.entries() / (double)GV.n : 0.0;
  res->n_assignments = nAssignAll;
  // ---- FinalizeReadAssignments: equivalence classes + missing coverage
  double tc = now_ms();
  PhaseTimer pt;
  EquivalenceClasses EC;
  TailDevice tail;
  bool tailOnDevice = false;
  if (getenv("T1K_HOST_TAIL") == nullptr && GV.n > 0) {
    if (int rc = device_tail_classes(ref, GV, nA, shards.threads(), tail, EC, &tailOnDevice)) return rc;
  }
  if (!tailOnDevice) {
    if (compactOnDevice) {           // the host path needs the allele ids after all
      compactGroups.allele.resize((size_t)std::max<int64_t>(GV.entries(), 1));
      CK(cudaMemcpy(compactGroups.allele.data(), dGAllele.p, (size_t)GV.entries() * 4, cudaMemcpyDeviceToHost));
      GV.allele = compactGroups.allele.data(); GV.dAllele = nullptr;
    }
    EC = EquivalenceClasses(); EC.build(GV, nA, shards.threads());
  }
  pt.lap(tailOnDevice ? "equivalence classes (device)" : "equivalence classes");
  if (tailOnDevice) res->n_launches += 3;
  res->n_groups = GV.n; res->n_ec = EC.size(); res->n_alleles = nA;
  if (res->missing_coverage) { if (int rc = t1k_missing_coverage(ref, res->missing_coverage)) return rc; }
  pt.lap("missing coverage");
  if (res->equivalent_class) memcpy(res->equivalent_class, EC.alleleEc.data(), (size_t)nA * 4);
  if (res->ec_allele_ptr && res->ec_alleles) {
    memcpy(res->ec_allele_ptr, EC.ecPtr.data(), EC.ecPtr.size() * 4);
    if (!EC.ecAlleles.empty()) memcpy(res->ec_alleles, EC.ecAlleles.data(), EC.ecAlleles.size() * 4);
  }
  res->ms_coalesce += (float)(now_ms() - tc);
  // ---- EM
  double te = now_ms();
  res->em_iterations = 0;
  if (res->abundance) memset(res->abundance, 0, (size_t)nA * 8);
  if (res->ec_abundance) memset(res->ec_abundance, 0, (size_t)nA * 8);
  if (EC.size() > 0) {
    EmInputs in;
    EmDevice emDev;
    if (tailOnDevice) {
      if (int rc = device_tail_matrix(ref, tail, EC, nA, comm, emDev)) return rc;
      in.build_vectors(GV, EC, prm->effective_len, prm->seq_weight, shards.threads());
      in.rowPtr.swap(tail.hRowPtr);
      res->n_launches += 6;
    } else in.build(GV, EC, prm->effective_len, prm->seq_weight, shards.threads());
    pt.lap(tailOnDevice ? "EM inputs (device)" : "EM inputs");
    if (pt.on) fprintf(stderr, "[t1k timing] groups %d entries %lld ECs %d nnz %lld\n", GV.n, (long long)GV.entries(), EC.size(), (long long)in.rowPtr[GV.n]);
    T1KEmProblem ep;
    memset(&ep, 0, sizeof(ep));
    ep.n_groups = GV.n; ep.n_ec = EC.size();
    ep.row_ptr = in.rowPtr.data(); ep.col = tailOnDevice ? nullptr : in.col.data(); ep.count = in.count.data(); ep.ec_len = in.ecLen.data(); ep.x0 = in.x0.data();
    ep.min_squarem_alpha = prm->min_squarem_alpha; ep.filter_frac = prm->filter_frac; ep.fast_sums = prm->em_fast_sums;
    ep.comm = comm;
    // the columns of the matrix are the group lists of the class representatives: handed over as they lie in EC
    std::vector<int64_t> colBeg, colEnd;
    if (!tailOnDevice) {
      colBeg.resize((size_t)EC.size()); colEnd.resize((size_t)EC.size());
      for (int32_t e = 0; e < EC.size(); ++e) { const int32_t rep = EC.ecAlleles[EC.ecPtr[e]]; colBeg[e] = EC.inPtr[rep]; colEnd[e] = EC.inPtr[rep + 1]; }
    }
    const EmColumns emCols = {colBeg.data(), colEnd.data(), EC.in.data(), EC.in.size()};
    if (prm->allele_major && prm->allele_gene) {
      ep.n_alleles = nA; ep.n_major = prm->n_major; ep.n_gene = prm->n_gene;
      ep.ec_allele_ptr = EC.ecPtr.data(); ep.ec_alleles = EC.ecAlleles.data();
      ep.allele_major = prm->allele_major; ep.allele_gene = prm->allele_gene;
    }
    std::vector<double> x(EC.size()), rc(EC.size());
    T1KEmResult er; er.x = x.data(); er.ec_read_count = rc.data(); er.iterations = 0;
    g_emTrusted = true; g_emCols = tailOnDevice ? nullptr : &emCols; g_emDev = tailOnDevice ? &emDev : nullptr;
    const int rcode = t1k_em_run(&ep, &er, ref->device);
    g_emTrusted = false; g_emCols = nullptr; g_emDev = nullptr;
    if (rcode) return rcode;
    pt.lap("t1k_em_run");
    res->em_iterations = er.iterations;
    res->ms_em_kernel = er.ms_kernel; res->n_launches += er.n_launches;
    res->em_nnz = (uint64_t)in.rowPtr[GV.n]; res->em_updates = 3 * er.iterations;
    if (res->abundance && res->ec_abundance)
      set_allele_abundance(rc.data(), in.ecLen.data(), EC.ecPtr.data(), EC.ecAlleles.data(), EC.size(), nA, res->abundance, res->ec_abundance);
    if (res->ec_read_count) memcpy(res->ec_read_count, rc.data(), (size_t)EC.size() * 8);
  }
  res->ms_em = (float)(now_ms() - te);
  // ---- RemoveLowLikelihoodAlleleInEquivalentClass (Genotyper.cpp:647)
  if (res->allele_span) memcpy(res->allele_span, spans.data(), spans.size() * 4);
  if (res->allele_kept) {
    if (!res->ec_abundance) return fail(T1K_ERR_ARG, "t1k_genotype: allele_kept needs ec_abundance");
    ec_likelihood_filter(EC.ecPtr.data(), EC.ecAlleles.data(), EC.size(), nA, ref->len.data(), res->ec_abundance, spans.data(), res->allele_kept);
  }
  if (prm->groups_out) {
    T1KGroups *g = new T1KGroups;
    g->G.ptr.swap(groups.ptr); g->G.ent.swap(groups.ent); g->G.first.swap(groups.first); g->G.hashes.swap(groups.hashes);
    g->G.assignedFragments = groups.assignedFragments;
    *prm->groups_out = g;
  }
  return T1K_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// edit strings for the analyzer (SeqSet::AddOverlapAlignmentInfo)
static_assert(sizeof(OvIn) == sizeof(T1KOverlap), "OvIn mirrors T1KOverlap");

extern "C" {

int t1k_align_info_batch(T1KRef *ref, const char *bases, const uint64_t *off, const uint32_t *len, uint32_t n_reads, const uint32_t *read_idx,
                         const T1KOverlap *ov, uint32_t n_items, int32_t flags, uint64_t *align_ptr, int8_t **align, uint64_t *align_bytes,
                         T1KAlignInfoStats *stats) {
  if (!ref || !align || !align_bytes || (n_items > 0 && (!read_idx || !ov || !align_ptr)) || (n_reads > 0 && (!bases || !off || !len)))
    return fail(T1K_ERR_ARG, "t1k_align_info_batch: bad argument");
  *align = nullptr; *align_bytes = 0;
  if (stats) memset(stats, 0, sizeof(*stats));
  CK(cudaSetDevice(ref->device));
  cudaStream_t st = ref->stream;
  size_t total = 0; int maxLen = KMER;
  for (uint32_t i = 0; i < n_reads; ++i) {
    if (len[i] > T1K_MAX_READ_LEN) return fail(T1K_ERR_ARG, "read longer than T1K_MAX_READ_LEN");
    total = std::max(total, (size_t)(off[i] + len[i]));
    maxLen = std::max(maxLen, (int)len[i]);
  }
  // output slots: 16-byte aligned, lent + lenp + 2 bytes (every op consumes a base of at least one side; + the -1)
  std::vector<u64> slot(n_items);
  u64 bytes = 0;
  for (uint32_t i = 0; i < n_items; ++i) {
    const T1KOverlap &o = ov[i];
    if (o.seqIdx == -1) { slot[i] = ~0ull; align_ptr[i] = ~0ull; continue; }
    if (o.seqIdx < 0 || o.seqIdx >= ref->nAlleles || read_idx[i] >= n_reads || (o.strand != 1 && o.strand != -1) || o.seqStart < 0 ||
        o.seqEnd < o.seqStart - 1 || o.seqEnd >= ref->len[o.seqIdx] || o.readStart < 0 || o.readEnd < o.readStart - 1 ||
        o.readEnd >= (int32_t)len[read_idx[i]])
      return fail(T1K_ERR_ARG, "t1k_align_info_batch: overlap coordinates outside the allele / read");
    slot[i] = bytes; align_ptr[i] = bytes;
    bytes += ((u64)(o.seqEnd - o.seqStart + 1) + (u64)(o.readEnd - o.readStart + 1) + 2 + 15) & ~(u64)15;
  }
  PhaseTimer pt;
  stale("t1k_align_info_batch");
  if (int rc = setup_assign_launch(ref, maxLen)) return rc;
  int8_t *host = (int8_t *)malloc(std::max<u64>(bytes, 16));
  if (!host) return fail(T1K_ERR_ARG, "t1k_align_info_batch: out of host memory");
  struct HostGuard { int8_t *p; ~HostGuard() { free(p); } } hg{host};
  if (n_items == 0 || bytes == 0) { hg.p = nullptr; *align = host; return T1K_OK; }
  DevMem dBases, dOff, dLen, planes, len16, dOv, dIdx, dSlot, dOut;
  const int rwords = read_words(ref->scrLen);
  CK(dBases.alloc(total)); CK(dOff.alloc((size_t)n_reads * 8)); CK(dLen.alloc((size_t)n_reads * 4));
  CK(planes.alloc((size_t)n_reads * 4 * rwords * 8)); CK(len16.alloc((size_t)n_reads * 2));
  CK(dOv.alloc((size_t)n_items * sizeof(OvIn))); CK(dIdx.alloc((size_t)n_items * 4)); CK(dSlot.alloc((size_t)n_items * 8)); CK(dOut.alloc(bytes));
  CK(cudaMemcpyAsync(dBases.p, bases, total, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dOff.p, off, (size_t)n_reads * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dLen.p, len, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dOv.p, ov, (size_t)n_items * sizeof(OvIn), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dIdx.p, read_idx, (size_t)n_items * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dSlot.p, slot.data(), (size_t)n_items * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ref->errFlag.p, 0, sizeof(int), st));
  CK(cudaMemsetAsync(ref->stats.p, 0, 8 * sizeof(unsigned long long), st));
  k_pack_reads<<<(n_reads + 127) / 128, 128, 0, st>>>(dBases.as<char>(), dOff.as<u64>(), dLen.as<u32>(), n_reads, rwords, planes.as<u64>(), len16.as<u16>(),
                                                        ref->errFlag.as<int>());
  CK(cudaGetLastError());
  AlnInfoParams P;
  P.R = ref->R; P.planes = planes.as<u64>(); P.rwords = rwords; P.maxLen = ref->scrLen; P.len = len16.as<u16>();
  P.ov = dOv.as<OvIn>(); P.readIdx = dIdx.as<u32>(); P.nItems = n_items; P.slot = dSlot.as<u64>(); P.out = dOut.as<u8>();
  P.laneScratch = ref->laneScratch.as<u8>(); P.err = ref->errFlag.as<int>(); P.stats = ref->stats.as<unsigned long long>();
  P.noDiag = (flags & 1) ? 1 : 0;
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
  CK(cudaEventRecord(ev0, st));
  const int blocks = (int)std::min<size_t>(ref->gridBlocks, ((size_t)n_items + WARPS_PER_BLOCK * 32 - 1) / (WARPS_PER_BLOCK * 32));
  k_align_info<<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(P);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ev1, st));
  int err = 0; unsigned long long hs[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(&err, ref->errFlag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hs, ref->stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(host, dOut.p, bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  pt.lap("align info: kernel + D2H");
  if (err & (ERR_READ_LEN | ERR_READ_CHAR)) return fail(T1K_ERR_ARG, "read contains a character outside ACGTN or is too long");
  if (err) return fail(T1K_ERR_UNSUPPORTED, "t1k_align_info_batch:" + decode_err(err));
  if (stats) {
    stats->n_diagonal = hs[0]; stats->n_dp = hs[1]; stats->dp_cells = hs[2];
    CK(cudaEventElapsedTime(&stats->ms_kernel, ev0, ev1));
  }
  hg.p = nullptr;
  *align = host; *align_bytes = bytes;
  return T1K_OK;
}

int t1k_dpx_peak(int32_t device, double *gops) {
  if (!gops) return fail(T1K_ERR_ARG, "t1k_dpx_peak: bad argument");
  int dev = 0;
  if (int rc = pick_device(device, &dev)) return rc;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  DevMem sink;
  CK(sink.alloc(4));
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {          // first round warms up
    CK(cudaEventRecord(ev0, 0));
    k_dpx_peak<<<blocks, threads>>>(iters, 12345 + rep, sink.as<int>());
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev1, 0));
    CK(cudaEventSynchronize(ev1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, ev0, ev1));
    if (rep > 0 && ms > 0) best = std::max(best, (double)blocks * threads * (double)iters * 8.0 / (ms * 1e-3) / 1e9);
  }
  *gops = best;
  return T1K_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// candidate filter of fastq-extractor
struct T1KFilter {
  int device = 0, nSM = 0, k = 0, hitLenReq = 0;
  double sim = 0.8;
  DevMem kinfo, entries, present, workCtr, errFlag, stats, hitBuf;
  cudaStream_t stream = nullptr;
  int gridBlocks = 0, seedCap = 0, rwords = 0, hitCap = 4096;
  ~T1KFilter() { if (stream) cudaStreamDestroy(stream); }
};

extern "C" {

int t1k_filter_create(const T1KFilterDesc *d, T1KFilter **out) {
  if (!d || !out || d->n_seqs <= 0 || !d->bases || !d->offset) return fail(T1K_ERR_ARG, "t1k_filter_create: bad argument");
  *out = nullptr;
  int dev;
  if (int rc = pick_device(d->device, &dev)) return rc;
  CK(cudaSetDevice(dev));
  int k = d->kmer_length;
  if (k == 0) {                      // SeqSet::InferKmerLength (SeqSet.hpp:2830-2845) under the extractor's floor of 9
    int64_t tot = d->offset[d->n_seqs] - d->offset[0];
    int digits = 0;
    while (tot) { ++digits; tot /= 4; }
    k = std::max(9, digits + 1);
  }
  if (k < 1 || k > 15) return fail(T1K_ERR_UNSUPPORTED, "candidate filter: k-mer length outside 1..15");
  std::vector<KmerInfo> kinfo;
  std::vector<KmerEntry> entries;
  if (!build_filter_index(d->n_seqs, d->bases, d->offset, k, kinfo, entries)) return fail(T1K_ERR_ARG, "reference contains a character outside ACGTN");
  for (size_t i = 0; i < entries.size(); ++i)
    if (entries[i].off >= (1u << 22)) return fail(T1K_ERR_UNSUPPORTED, "sequence longer than 2^22 bases");
  entries.resize(entries.size() + 4, KmerEntry{0xffffffffu, 0, 0, 0});
  std::vector<u32> present((kinfo.size() + 31) / 32, 0);
  for (size_t c = 0; c + 1 < kinfo.size(); ++c) if (kinfo[c + 1].pstart > kinfo[c].pstart) present[c >> 5] |= 1u << (c & 31);
  T1KFilter *f = new T1KFilter;
  f->device = dev; f->k = k; f->hitLenReq = d->hit_len_required; f->sim = d->similarity;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e == cudaSuccess) { f->nSM = prop.multiProcessorCount; e = f->kinfo.alloc(kinfo.size() * sizeof(KmerInfo)); }
  if (e == cudaSuccess) e = cudaMemcpy(f->kinfo.p, kinfo.data(), kinfo.size() * sizeof(KmerInfo), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = f->entries.alloc(entries.size() * sizeof(KmerEntry));
  if (e == cudaSuccess) e = cudaMemcpy(f->entries.p, entries.data(), entries.size() * sizeof(KmerEntry), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = f->present.alloc(present.size() * 4);
  if (e == cudaSuccess) e = cudaMemcpy(f->present.p, present.data(), present.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = f->workCtr.alloc(4);
  if (e == cudaSuccess) e = f->errFlag.alloc(4);
  if (e == cudaSuccess) e = f->stats.alloc(3 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete f; return fail(T1K_ERR_CUDA, std::string("t1k_filter_create: ") + cudaGetErrorString(e)); }
  *out = f;
  return T1K_OK;
}

void t1k_filter_destroy(T1KFilter *f) {
  if (!f) return;
  cudaSetDevice(f->device);
  delete f;
}

int t1k_filter_batch(T1KFilter *f, const char *bases, const uint64_t *off, const uint32_t *len, uint32_t n, uint8_t *good, T1KFilterStats *stats) {
  if (!f || (n > 0 && (!bases || !off || !len || !good))) return fail(T1K_ERR_ARG, "t1k_filter_batch: bad argument");
  CK(cudaSetDevice(f->device));
  stale("t1k_filter_batch");
  if (stats) { stats->windows = stats->entries = stats->chained = 0; stats->ms_kernel = 0; stats->kmer_length = f->k; }
  if (n == 0) return T1K_OK;
  cudaStream_t st = f->stream;
  size_t total = 0; int maxLen = 1;
  for (uint32_t i = 0; i < n; ++i) {
    if (len[i] > T1K_MAX_READ_LEN) return fail(T1K_ERR_ARG, "read longer than T1K_MAX_READ_LEN");
    total = std::max(total, (size_t)(off[i] + len[i]));
    maxLen = std::max(maxLen, (int)len[i]);
  }
  const int rwords = read_words(maxLen <= 255 ? 255 : MAX_READ_LEN);
  const int seedCap = std::max(64, (maxLen + 31) & ~31);
  const size_t smem = filter_smem_bytes(seedCap, rwords) * WARPS_PER_BLOCK;
  if (!f->gridBlocks || seedCap > f->seedCap || rwords != f->rwords) {
    CK(cudaFuncSetAttribute((const void *)k_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSM = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, (const void *)k_filter, WARPS_PER_BLOCK * 32, smem));
    if (perSM < 1) return fail(T1K_ERR_UNSUPPORTED, "k_filter does not fit on an SM");
    f->gridBlocks = perSM * f->nSM; f->seedCap = seedCap; f->rwords = rwords;
    CK(f->hitBuf.alloc((size_t)f->gridBlocks * WARPS_PER_BLOCK * ((size_t)4 * f->hitCap + filt::FILTER_USED_BYTES / 4) * sizeof(u32)));
  }
  DevMem dBases, dOff, dLen, planes, len16, dGood;
  CK(dBases.alloc(total)); CK(dOff.alloc((size_t)n * 8)); CK(dLen.alloc((size_t)n * 4));
  CK(planes.alloc((size_t)n * 4 * rwords * 8)); CK(len16.alloc((size_t)n * 2)); CK(dGood.alloc(n));
  CK(cudaMemcpyAsync(dBases.p, bases, total, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dOff.p, off, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dLen.p, len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(f->errFlag.p, 0, 4, st));
  CK(cudaMemsetAsync(f->workCtr.p, 0, 4, st));
  CK(cudaMemsetAsync(f->stats.p, 0, 3 * sizeof(unsigned long long), st));
  k_pack_reads<<<(n + 127) / 128, 128, 0, st>>>(dBases.as<char>(), dOff.as<u64>(), dLen.as<u32>(), n, rwords, planes.as<u64>(), len16.as<u16>(), f->errFlag.as<int>());
  CK(cudaGetLastError());
  FilterParams P;
  P.kinfo = f->kinfo.as<KmerInfo>(); P.entries = f->entries.as<KmerEntry>(); P.present = f->present.as<u32>(); P.k = f->k; P.hitLenReq = f->hitLenReq; P.sim = f->sim;
  P.planes = planes.as<u64>(); P.rwords = rwords; P.len = len16.as<u16>(); P.nReads = n; P.good = dGood.as<u8>();
  P.hitBuf = f->hitBuf.as<u32>(); P.hitCap = f->hitCap; P.seedCap = f->seedCap; P.workCtr = f->workCtr.as<unsigned int>();
  P.err = f->errFlag.as<int>(); P.stats = f->stats.as<unsigned long long>();
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
  CK(cudaEventRecord(ev0, st));
  k_filter<<<f->gridBlocks, WARPS_PER_BLOCK * 32, filter_smem_bytes(f->seedCap, rwords) * WARPS_PER_BLOCK, st>>>(P);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ev1, st));
  int err = 0;
  unsigned long long hs[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(good, dGood.p, n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&err, f->errFlag.p, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hs, f->stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (err & (ERR_READ_LEN | ERR_READ_CHAR)) return fail(T1K_ERR_ARG, "read contains a character outside ACGTN or is too long");
  if (err) return fail(T1K_ERR_UNSUPPORTED, "t1k_filter_batch:" + decode_err(err));
  if (stats) {
    stats->windows = hs[0]; stats->entries = hs[1]; stats->chained = hs[2];
    CK(cudaEventElapsedTime(&stats->ms_kernel, ev0, ev1));
  }
  return T1K_OK;
}

}  // extern "C"
