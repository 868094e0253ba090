/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of T1K's genotyping hot path, written from the reference's
 * behaviour (file:line citations in t1k_oracle.cpp), used exclusively by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for the CUDA path.
 * The product (t1k_b200/csrc) never includes, links or calls anything declared here.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path,
 * so the restatement is pinned by differential testing against the reference itself compiled
 * in-container (oracle/_ref/ref_harness, built from /root/reference by oracle/Makefile):
 * tests/test_oracle_vs_reference.py (runs where /root/reference exists) and the committed
 * outputs of that harness under tests/golden/ (run everywhere).
 */
#ifndef T1K_ORACLE_H
#define T1K_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct T1KOracle T1KOracle;

/* One (read-end, allele) alignment; field meaning = struct _overlap, SeqSet.hpp:89-101.
 * similarity is derived: matchCnt / (readSpan + seqSpan + 2*leftClip + 2*rightClip). */
typedef struct {
  int32_t seqIdx, readStart, readEnd, seqStart, seqEnd, strand, matchCnt, relaxedMatchCnt, leftClip, rightClip;
} OracleOverlap;

/* One (fragment, allele) assignment = struct _readAssignment, Genotyper.hpp:44-56 */
typedef struct {
  int32_t alleleIdx, start, end;
  float weight, qual, adjustWeight;
} OracleAssignment;

/* alleles: concatenated upper-case ACGTN; off[n+1]; exons: exonPtr[n+1] into exonSE pairs (inclusive). */
T1KOracle *t1ko_create(int32_t nAlleles, const char *bases, const int64_t *off, const int32_t *exonPtr,
                       const int32_t *exonSE, const int32_t *seqWeight, double similarity, int32_t relaxIntron);
void t1ko_destroy(T1KOracle *o);

/* ---- candidate filter of fastq-extractor: FastqExtractor.cpp:89-119, SeqSet::HasHitInSet SeqSet.hpp:1915-1990.
 * The reference set is loaded as InputRefFa does (every record, no collapsing); k and hitLenRequired as main() derives them. */
int32_t t1ko_infer_kmer_length(int64_t totalLength);
T1KOracle *t1ko_filter_create(int32_t nSeqs, const char *bases, const int64_t *off, int32_t k, int32_t hitLenRequired, double similarity);
int32_t t1ko_is_low_complexity(const char *read);
int32_t t1ko_has_hit_in_set(T1KOracle *o, const char *read);
int32_t t1ko_is_good_candidate(T1KOracle *o, const char *read);

/* AlignAlgo::GlobalAlignment.  ops receives 0 M,1 X,2 I,3 D; returns score, *nOps set. */
int32_t t1ko_global_alignment(const char *t, int32_t lent, const char *p, int32_t lenp, int8_t *ops, int32_t *nOps);

/* SeqSet::AssignRead.  Returns the count (or -1 exactly where the reference returns -1); writes up to cap records. */
int32_t t1ko_assign_read(T1KOracle *o, const char *read, int32_t weight, OracleOverlap *out, int32_t cap);

/* coverage of count[consensus base] per position of one allele */
void t1ko_coverage(T1KOracle *o, int32_t allele, int32_t *out);
void t1ko_coverage_reset(T1KOracle *o);
int32_t t1ko_allele_len(T1KOracle *o, int32_t allele);
int32_t t1ko_effective_len(T1KOracle *o, int32_t allele);
/* SeqSet::GetSeqMissingBaseCoverage(allele, 0.01) */
int32_t t1ko_missing_coverage(T1KOracle *o, int32_t allele);

/* SeqSet::ReadAssignmentToFragmentAssignment + Genotyper::SetReadAssignments.
 * o2 == NULL means single-end.  Returns number of assignments written (<= cap). */
int32_t t1ko_fragment_assign(T1KOracle *o, const OracleOverlap *o1, int32_t n1, const OracleOverlap *o2, int32_t n2,
                             int32_t hasN, int32_t maxAssign, OracleAssignment *out, int32_t cap);

/* SQUAREM EM over read groups x equivalence classes (Genotyper.hpp:372-437,1234-1316).
 * rowPtr[G+1], col[nnz] EC ids, count[G], ecLen[E], x0[E] in; x[E], ecReadCount[E] out.
 * maskMajor/maskGene give, per EC member allele list (ecAllelePtr/ecAlleles), what the every-10 mask needs:
 * alleleMajor[nAlleles], alleleGene[nAlleles]; pass nAlleles = 0 to disable masking. Returns iterations. */
int32_t t1ko_em(int32_t G, int32_t E, const int64_t *rowPtr, const int32_t *col, const double *count,
                const int32_t *ecLen, const double *x0, double minAlpha, double filterFrac,
                int32_t nAlleles, const int32_t *ecAllelePtr, const int32_t *ecAlleles,
                const int32_t *alleleMajor, const int32_t *alleleGene, int32_t nMajor, int32_t nGene,
                double *xOut, double *ecReadCountOut);

#ifdef __cplusplus
}
#endif
#endif
