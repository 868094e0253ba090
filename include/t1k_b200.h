/* t1k_b200 — C ABI of the B200-native T1K genotyping hot path (align + EM).
 *
 * The reference (mourisl/T1K) has no plugin/FFI layer: the seam is the set of C++ member calls that
 * Genotyper.cpp / Analyzer.cpp make into SeqSet and Genotyper.  Each entry point
 * below replaces one of those call sites; the citation names the reference interface it stands for
 * (paths relative to the reference checkout).  Plain pointers and sizes only; every function returns
 * an int status (0 = ok) and never aborts the process.  All device work is hand-written sm_100a CUDA;
 * there is no CPU fallback: without a CUDA device every compute entry point returns T1K_ERR_NO_DEVICE.
 */
#ifndef T1K_B200_H
#define T1K_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  T1K_OK = 0,
  T1K_ERR_NO_DEVICE = 1,    /* no usable CUDA device */
  T1K_ERR_CUDA = 2,         /* a CUDA runtime call failed; t1k_last_error() has the text */
  T1K_ERR_ARG = 3,          /* bad argument (NULL, read longer than T1K_MAX_READ_LEN, non-ACGTN base, ...) */
  T1K_ERR_UNSUPPORTED = 4,  /* input outside the supported envelope (band wider than T1K_MAX_BAND, scratch overflow) */
  T1K_ERR_NCCL = 5
};

#define T1K_MAX_READ_LEN 1000
#define T1K_KMER 11

typedef struct T1KRef T1KRef;               /* allele reference + k-mer index + coverage, resident in HBM */
typedef struct T1KAssignment T1KAssignment; /* per-read-end overlap lists, resident in HBM */
typedef struct T1KComm T1KComm;             /* NCCL communicator of the read-sharded path: one rank per process/GPU */
typedef struct T1KGroups T1KGroups;         /* coalesced read groups (host): the rows of the EM's incidence matrix */

/* One (read-end, allele) alignment: the result fields of `struct _overlap` (SeqSet.hpp:89-101).
 * similarity = matchCnt / (readEnd-readStart+1 + seqEnd-seqStart+1 + 2*leftClip + 2*rightClip)
 * exactly as SeqSet.hpp:2066-2067,2085-2086 compute it (recompute in double on the host). */
typedef struct {
  int32_t seqIdx, readStart, readEnd, seqStart, seqEnd, strand, matchCnt, relaxedMatchCnt, leftClip, rightClip;
} T1KOverlap;

/* One (fragment, allele) entry of the sparse read x allele matrix: `struct _readAssignment`
 * (Genotyper.hpp:44-56). */
typedef struct {
  int32_t alleleIdx, start, end;
  float weight, qual, adjustWeight;
} T1KReadAssignment;

/* Reference description.  Replaces Genotyper::InitRefSet -> SeqSet::InputRefSeq (Genotyper.hpp:707-730,
 * SeqSet.hpp:906-982) for an already de-duplicated allele list (identical sequences collapsed by the
 * caller, weight++ — Genotyper.hpp:717-725). */
typedef struct {
  int32_t n_alleles;
  const char *bases;        /* concatenated upper-case ACGTN */
  const int64_t *offset;    /* [n_alleles+1] */
  const int32_t *exon_ptr;  /* [n_alleles+1] into exon_se pairs */
  const int32_t *exon_se;   /* (start,end) inclusive, SeqSet.hpp:933-976 */
  double similarity;        /* -s, SeqSet::SetRefSeqSimilarity (Genotyper.cpp:350) */
  int32_t relax_intron;     /* --relaxIntronAlign, SeqSet::SetRelaxIntronAlign (Genotyper.cpp:351) */
  int32_t device;           /* CUDA device ordinal, -1 = current */
} T1KRefDesc;

const char *t1k_last_error(void);
int t1k_device_count(int *count);

int t1k_ref_create(const T1KRefDesc *desc, T1KRef **out);
void t1k_ref_destroy(T1KRef *ref);
int t1k_ref_n_alleles(const T1KRef *ref);

/* SeqSet::AssignRead for a batch of UNIQUE read-ends (Genotyper.cpp:149,472; Analyzer.cpp:142,476).
 * bases/off/len: concatenated reads (host memory, pinned or not); weight[i] = number of duplicates
 * (>=1 adds base coverage, 0 = analyzer mode, SeqSet.hpp:2253).  The result stays on the device. */
int t1k_assign_batch(T1KRef *ref, const char *bases, const uint64_t *off, const uint32_t *len,
                     const int32_t *weight, uint32_t n_reads, T1KAssignment **out);
void t1k_assignment_destroy(T1KAssignment *a);
/* Asynchronous variant ("async variant with stream / double-buffered pinned staging").  Returns at once; the batch
 * is aligned on the reference's stream by a worker of the library while the caller de-duplicates and stages the next batch —
 * what the pthread fan-out of Genotyper.cpp:481-507 gives the reference.  Jobs of one T1KRef run in submission order (coverage
 * and results are those of the same sequence of synchronous calls).  The input arrays must stay valid and unchanged until
 * t1k_assign_wait returns; t1k_pinned_alloc gives page-locked staging buffers so that the upload is a true DMA.
 * t1k_assign_wait blocks until the job is done, hands out its result and status (error text in t1k_last_error) and frees the job. */
typedef struct T1KAssignJob T1KAssignJob;
int t1k_assign_batch_async(T1KRef *ref, const char *bases, const uint64_t *off, const uint32_t *len, const int32_t *weight,
                           uint32_t n_reads, T1KAssignJob **job);
int t1k_assign_wait(T1KAssignJob *job, T1KAssignment **out);
int t1k_pinned_alloc(uint64_t bytes, void **p);
void t1k_pinned_free(void *p);
/* Host copy, per read in the reference's output order (`assign` of SeqSet.hpp:2300):
 * row_ptr[n_reads+1] (caller-allocated), ret[n_reads] = AssignRead's return value (count or -1),
 * records: call once with records==NULL to get *total, then with a buffer of *total entries. */
int t1k_assignment_fetch(T1KAssignment *a, uint64_t *row_ptr, int32_t *ret, T1KOverlap *records, uint64_t *total);

/* Work counters of one t1k_assign_batch call, for roofline accounting (no reference counterpart). */
typedef struct {
  uint64_t postings, candidates, tiles, records;   /* postings read, seed overlaps chained, allele tiles, records kept */
  float ms_kernel;                                 /* device time of the AssignRead kernels (CUDA events) */
  int32_t grid_blocks, hit_cap, n_sm;              /* launch geometry */
} T1KAssignStats;
int t1k_assignment_stats(const T1KAssignment *a, T1KAssignStats *out);

/* posWeight[].count[consensus base] per base of every allele, concatenated by `offset` (Q11:
 * the only counter GetSeqMissingBaseCoverage reads, SeqSet.hpp:2727-2731). */
int t1k_coverage_fetch(T1KRef *ref, int32_t *out /* [offset[n_alleles]] */);
int t1k_coverage_reset(T1KRef *ref);
/* SeqSet::GetSeqMissingBaseCoverage(i, 0.01) for every allele (SeqSet.hpp:2717-2755). */
int t1k_missing_coverage(T1KRef *ref, int32_t *out /* [n_alleles] */);

/* SeqSet::ReadAssignmentToFragmentAssignment + Genotyper::SetReadAssignments for a batch of fragments
 * (Genotyper.cpp:183-187,542-554; SeqSet.hpp:2310-2655; Genotyper.hpp:778-832).
 * end1[i]/end2[i] index read-ends of `a` (end2 == NULL: single-end); has_n[i] = either mate holds an N.
 * Output CSR (library-allocated host memory, free with t1k_free): row_ptr[n_frag+1], entries in the
 * reference's order. */
int t1k_pair_batch(T1KRef *ref, T1KAssignment *a, const uint32_t *end1, const uint32_t *end2,
                   const uint8_t *has_n, uint32_t n_frag, int32_t max_assign,
                   uint64_t **row_ptr, T1KReadAssignment **entries,
                   uint8_t *fragment_assigned /* [n_frag] or NULL: fragmentAssignment.size() > 0, Genotyper.cpp:564 */);
void t1k_free(void *p);

/* Genotyper::QuantifyAlleleEquivalentClass main loop (Genotyper.hpp:1234-1316) on the device.
 * Rows = read groups, columns = allele equivalence classes. */
typedef struct {
  int32_t n_groups, n_ec;
  const int64_t *row_ptr;      /* [n_groups+1] */
  const int32_t *col;          /* EC ids, per group in first-appearance order (Genotyper.hpp:1165-1189) */
  const double *count;         /* readGroupInfo[].count */
  const int32_t *ec_len;       /* ecInfo[].length */
  const double *x0;            /* initial ecAbundance0 (sum of member seqWeight) */
  double min_squarem_alpha;    /* --squaremMinAlpha, 0 = unset */
  double filter_frac;          /* --frac */
  /* every-10-iterations mask (Genotyper.hpp:1292-1313); n_alleles = 0 disables it */
  int32_t n_alleles, n_major, n_gene;
  const int32_t *ec_allele_ptr, *ec_alleles;   /* members per EC (first = representative) */
  const int32_t *allele_major, *allele_gene;   /* [n_alleles] */
  /* 0: every sum in the reference's sequential order, separately rounded (abundances bit-identical to the
   *    reference's x86 result; the dependent add chains are serial);
   * 1: warp/block tree reductions (fixed order, run-to-run reproducible, ~1e-16 relative from the reference). */
  int32_t fast_sums;
  /* read-sharded EM: every rank passes the SAME problem; rank r runs the E-step over its contiguous
   * row range (t1k_em_partition) and one ncclAllReduce(sum, f64, n_ec) per EMupdate combines ecReadCount; the
   * M-step / SQUAREM vector steps are replicated.  NULL = single GPU.  With more than one rank the sums no longer
   * run in the reference's order (results within 1e-5 relative, deterministic for a given world size). */
  T1KComm *comm;
} T1KEmProblem;

typedef struct {
  double *x;              /* [n_ec] final ecAbundance0 */
  double *ec_read_count;  /* [n_ec] */
  int32_t iterations;
  float ms_kernel;        /* device time of the whole EM loop (CUDA events) */
  uint64_t n_launches;    /* kernels launched */
} T1KEmResult;

int t1k_em_run(const T1KEmProblem *p, T1KEmResult *r, int32_t device);

/* The whole hot path for one sample: de-duplicate read-ends, align, pair, coalesce read groups,
 * build equivalence classes, EM (Genotyper.cpp:450-646).  reads are fixed-stride host buffers
 * (stride bytes per read, '\0'-padded); reads2 == NULL for single-end. */
typedef struct {
  int32_t max_assign;          /* -n (2000) */
  double min_squarem_alpha, filter_frac;
  const int32_t *seq_weight;   /* [n_alleles] SeqSet::GetSeqWeight */
  const int32_t *effective_len;/* [n_alleles] SeqSet::GetSeqEffectiveLen (after InitAlleleInfo's adjustment) */
  const int32_t *allele_major, *allele_gene; int32_t n_major, n_gene;
  int32_t em_fast_sums;        /* T1KEmProblem.fast_sums */
  /* read-sharded run: reads1/reads2 are THIS rank's fragments.  Alignment + pairing + coalescing run per rank with no
   * data-path collective; then a status exchange (a failure on one rank fails every rank instead of hanging the others),
   * one int32 all-reduce of the base coverage, the read-group tables merged 1/world per rank (all-to-all of the hash
   * partitions, all-gather of the merged partitions) and the sharded EM.  Per-allele outputs are identical on all ranks;
   * fragment_assigned / n_unique_ends / n_overlaps / timings are this rank's.  NULL = single GPU. */
  T1KComm *comm;
  /* optional: the coalesced read groups (Genotyper::readAssignments after CoalesceReadAssignments) as a T1KGroups handle for
   * t1k_groups_fetch / t1k_groups_destroy — what a driver needs to continue with the reference's own
   * FinalizeReadAssignments / allele selection.  NULL = not wanted. */
  T1KGroups **groups_out;
} T1KGenotypeParams;

typedef struct {
  int32_t n_alleles;
  double *abundance, *ec_abundance;   /* [n_alleles] alleleInfo[].abundance / .ecAbundance */
  int32_t *equivalent_class;          /* [n_alleles] (-1: no reads) */
  int32_t *missing_coverage;          /* [n_alleles] */
  uint8_t *fragment_assigned;         /* [n_frag] */
  int32_t em_iterations, n_groups, n_ec, assigned_fragments;
  uint64_t n_unique_ends, n_overlaps, n_assignments;
  double avg_alleles_per_read;        /* Genotyper::GetAverageReadAssignmentCnt: over read groups */
  float ms_dedup, ms_align, ms_pair, ms_coalesce, ms_em;   /* wall time of the phases of this call (host clock) */
  float ms_align_kernel, ms_pair_kernel, ms_em_kernel;      /* device time (CUDA events) inside the AssignRead kernels / k_pair / the EM kernels */
  uint64_t n_postings, n_candidates;                        /* k-mer postings read and seed overlaps chained (roofline accounting) */
  uint64_t n_launches;                                      /* kernels launched by this call */
  float ms_prep_wait;                                       /* part of ms_dedup the device stage had to wait for (not overlapped) */
  float ms_exchange;                                        /* read-sharded run: status / table exchange + merge (inside ms_coalesce) */
  uint64_t n_pair_records;                                  /* overlap records of both mates summed over the fragments (k_pair roofline) */
  uint64_t em_nnz;                                          /* non-zeros of the read-group x EC matrix */
  int32_t em_updates;                                       /* EMupdate calls (3 per SQUAREM iteration) */
  double *ec_read_count;                                    /* optional, caller-allocated [n_alleles]: ecReadCount of EC e at [e] (the argument
                                                               of Genotyper::SetAlleleAbundance, Genotyper.hpp:957) */
  int32_t *ec_allele_ptr, *ec_alleles;                      /* optional, caller-allocated [n_alleles+1] / [n_alleles]: members of every EC in the
                                                               order of Genotyper::equivalentClassToAlleles (Genotyper.hpp:1072-1139) */
  /* SURVEY.md §8f N4 — Genotyper::RemoveLowLikelihoodAlleleInEquivalentClass (Genotyper.hpp:1371-1460, called at Genotyper.cpp:647
   * right after the EM).  Optional, caller-allocated: allele_kept[n_alleles] = 1 when the allele stays in its equivalence
   * class (likelihood pow(covered range / length, ecAbundance) within 0.05 of the class's best), 0 when it is removed or has no
   * class; allele_span[2 * n_alleles] = the covered range the likelihood is computed from (min start at [a], INT32_MAX without
   * entries; max end at [n_alleles + a], -1 without).  Read-sharded runs reduce the ranges over the ranks (one int32 all-reduce). */
  uint8_t *allele_kept;
  int32_t *allele_span;
} T1KGenotypeResult;

int t1k_genotype(T1KRef *ref, const char *reads1, const char *reads2, uint32_t stride, uint32_t n_frag,
                 const T1KGenotypeParams *params, T1KGenotypeResult *res /* caller-allocated arrays */);

/* ---- the candidate filter of fastq-extractor, the stage in front of the hot path.
 * IsGoodCandidate (FastqExtractor.cpp:113-118) = !IsLowComplexity (:89-112) && SeqSet::HasHitInSet (SeqSet.hpp:1915-1990) for a
 * batch of reads against the sequences of the extraction reference.  kmer_length / hit_len_required are what
 * FastqExtractor.cpp:381-418 sets up (k = max(9, SeqSet::InferKmerLength(total length)); hitLenRequired from the mean read
 * length); kmer_length = 0 infers k from the total length as the reference does. */
typedef struct T1KFilter T1KFilter;
typedef struct {
  int32_t n_seqs;
  const char *bases;        /* concatenated upper-case ACGTN */
  const int64_t *offset;    /* [n_seqs+1] */
  int32_t kmer_length;      /* 0 = max(9, InferKmerLength) */
  int32_t hit_len_required;
  double similarity;        /* -s of fastq-extractor (SeqSet::SetRefSeqSimilarity) */
  int32_t device;           /* CUDA device ordinal, -1 = current */
} T1KFilterDesc;
typedef struct {
  uint64_t windows, entries, chained;   /* k-mer windows looked up, index entries swept, reads that reached the chaining */
  float ms_kernel;                      /* device time of k_filter (CUDA events) */
  int32_t kmer_length;
} T1KFilterStats;
int t1k_filter_create(const T1KFilterDesc *desc, T1KFilter **out);
void t1k_filter_destroy(T1KFilter *f);
/* good[i] = IsGoodCandidate(read i); a read pair is kept when either mate is good (FastqExtractor.cpp:199-212).
 * stats may be NULL. */
int t1k_filter_batch(T1KFilter *f, const char *bases, const uint64_t *off, const uint32_t *len, uint32_t n_reads,
                     uint8_t *good, T1KFilterStats *stats);

/* ---- reads into memory as Genotyper.cpp:363-454 gets them from ReadFiles / kseq (ReadFiles.hpp:155-204):
 * FASTA or FASTQ, plain or gzip; path2 == NULL for single-end.  Sequences only, exactly as they stand in the files (kseq's record
 * grammar, see csrc/t1k_reads.hpp), laid out as the fixed-stride NUL-padded buffers t1k_genotype takes (stride = longest read + 1;
 * page-locked when a CUDA device is present so that the chunk uploads are DMAs).  Host-side; needs no device. */
typedef struct {
  char *reads1, *reads2;       /* [n_frag * stride]; reads2 NULL for single-end */
  uint32_t stride, n_frag, max_len;
  int32_t pinned;              /* internal: how the buffers were allocated */
} T1KReads;
int t1k_reads_load(const char *path1, const char *path2, T1KReads *out);
void t1k_reads_free(T1KReads *r);

/* ---- the analyzer, second caller of the boundary (Analyzer.cpp:467-669).  It calls AssignRead with weight 0
 * (t1k_assign_batch above), pairs on the host over the fetched records, and after the EM asks for the edit string of every
 * overlap it kept: SeqSet::AddFragmentAlignmentInfo -> AddOverlapAlignmentInfo (SeqSet.hpp:2757-2778, 2657-2680) =
 * AlignAlgo::GlobalAlignment(allele[seqStart..seqEnd], strand-adjusted read[readStart..readEnd]) (AlignAlgo.hpp:215-421).
 * One call for a batch of (read, overlap) items: read_idx[i] indexes the reads (bases/off/len as in t1k_assign_batch),
 * ov[i] is a record of t1k_assignment_fetch (coordinates on the record's strand).  *align is library-allocated (t1k_free);
 * item i's string starts at *align + align_ptr[i] and ends with -1 exactly as `_overlap::align` does (0 EDIT_MATCH,
 * 1 EDIT_MISMATCH, 2 EDIT_INSERT, 3 EDIT_DELETE, AlignAlgo.hpp:7-10); align_ptr[i] = UINT64_MAX for seqIdx == -1 (the
 * reference leaves `align` unset, SeqSet.hpp:2659-2660).  flags bit 0: run the band DP for every item (no certified-diagonal
 * shortcut; results are identical, used for A/B checks and the DP roofline).  stats may be NULL. */
typedef struct {
  uint64_t n_diagonal, n_dp, dp_cells;   /* items written from the mismatch plane / by the band DP, band cells of the latter */
  float ms_kernel;                       /* device time of k_align_info (CUDA events) */
} T1KAlignInfoStats;
int t1k_align_info_batch(T1KRef *ref, const char *bases, const uint64_t *off, const uint32_t *len, uint32_t n_reads,
                         const uint32_t *read_idx, const T1KOverlap *ov, uint32_t n_items, int32_t flags,
                         uint64_t *align_ptr /* [n_items], caller-allocated */, int8_t **align, uint64_t *align_bytes,
                         T1KAlignInfoStats *stats);
/* Measured DPX issue rate of the device (G max(a+b,c) operations per second): the ceiling the band DP's cell rate is
 * reported against ("Banded DP").  No reference counterpart. */
int t1k_dpx_peak(int32_t device, double *gops);

/* ---- multi-GPU plumbing.  The launcher (torchrun + torch.distributed, MPI, a shared file ...) moves the 128-byte
 * unique id from rank 0 to the other ranks; everything on the data path is NCCL over NVLink. */
#define T1K_UNIQUE_ID_BYTES 128
int t1k_comm_unique_id(uint8_t *id /* [T1K_UNIQUE_ID_BYTES] */);
int t1k_comm_create(const uint8_t *id, int32_t rank, int32_t world, int32_t device, T1KComm **out);
void t1k_comm_destroy(T1KComm *comm);
/* sums the base coverage of `ref` over the ranks (in place; every rank ends with the total) */
int t1k_coverage_allreduce(T1KRef *ref, T1KComm *comm);

/* ---- host-side model steps kept in the reference's order (no device needed).
 * Genotyper::CoalesceReadAssignments (Genotyper.hpp:841-908): fragments with the same allele set merge into one
 * read group, float32 weights accumulate in fragment order.  Rows may be in any allele order. */
int t1k_groups_create(T1KGroups **out);
void t1k_groups_destroy(T1KGroups *g);
int t1k_groups_add_fragments(T1KGroups *g, const uint64_t *row_ptr, const T1KReadAssignment *entries, uint32_t n_frag);
/* the table as one relocatable blob (free with t1k_free) and the merge of another rank's blob into `g` */
int t1k_groups_serialize(const T1KGroups *g, void **blob, uint64_t *bytes);
int t1k_groups_merge(T1KGroups *g, const void *blob, uint64_t bytes);
/* n_groups / total entries / assigned fragments; then ptr[n_groups+1] and entries (caller-allocated, may be NULL) */
int t1k_groups_fetch(const T1KGroups *g, int32_t *n_groups, uint64_t *n_entries, uint64_t *assigned_fragments,
                     int64_t *ptr, T1KReadAssignment *entries);
/* Genotyper::RemoveLowLikelihoodAlleleInEquivalentClass (Genotyper.hpp:1371-1460) over a group table: ec_allele_ptr / ec_alleles
 * = Genotyper::equivalentClassToAlleles, allele_len = SeqSet::GetSeqConsensusLen, ec_abundance = alleleInfo[].ecAbundance
 * (T1KGenotypeResult.ec_abundance).  allele_kept[n_alleles], allele_span[2 * n_alleles] (may be NULL) as in T1KGenotypeResult. */
int t1k_groups_ec_filter(const T1KGroups *g, int32_t n_alleles, const int32_t *allele_len, const double *ec_abundance,
                         const int32_t *ec_allele_ptr, const int32_t *ec_alleles, int32_t n_ec, uint8_t *allele_kept, int32_t *allele_span);
/* contiguous row ranges of the EM problem balanced by non-zeros: bounds[world+1] */
int t1k_em_partition(const int64_t *row_ptr, int32_t n_groups, int32_t world, int32_t *bounds);

#ifdef __cplusplus
}
#endif
#endif
