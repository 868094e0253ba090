// Drop-in `genotyper` for run-t1k: the reference's stage-1 driver (Genotyper.cpp:194-738) with its three compute
// phases forwarded to the B200 library through the C ABI of include/t1k_b200.h:
//
//   PHASE A  SeqSet::AssignRead per unique read-end        (Genotyper.cpp:463-507)  -> t1k_assign_batch
//   PHASE B  ReadAssignmentToFragmentAssignment + SetReadAssignments (Genotyper.cpp:531-621) -> t1k_pair_batch
//   PHASE C  Genotyper::QuantifyAlleleEquivalentClass      (Genotyper.cpp:644)      -> t1k_em_run
//   base coverage read by FinalizeReadAssignments          (Genotyper.hpp:934)      <- t1k_coverage_fetch
//
// Everything else is the reference's own code, compiled from the reference checkout (-I$REF): FASTA/FASTQ parsing,
// InitRefSet / InitAlleleInfo, CoalesceReadAssignments, FinalizeReadAssignments, allele selection and the writers.
// This file is the binding a T1K maintainer would add; it contains no reference source, only calls into it.
// `private` is opened for the two classes instead of patching friend accessors into the reference headers.
//
// Build (needs the reference checkout):  make -C integration REF=/path/to/T1K
// Flags, outputs and log lines are those of the reference `genotyper`; `-t` is accepted and ignored for the phases
// that run on the GPU; T1K_DEVICE selects the CUDA device (default 0).
#include <getopt.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#define private public
#define protected public
#include "Genotyper.hpp"
#undef private
#undef protected

#include "../include/t1k_b200.h"

// defs.h externs (Genotyper.cpp:37-42 defines them in the reference's driver)
char nucToNum[26] = {0, -1, 1, -1, -1, -1, 2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1};
char numToNuc[4] = {'A', 'C', 'G', 'T'};

// same option letters and long names as the reference driver (Genotyper.cpp:13-57)
static const char usage[] =
    "genotyper (B200 hot path) -f ref.fa {-u reads | -1 reads_1 -2 reads_2} [options]\n"
    "  -a FILE   allele abundances to load instead of running the EM\n"
    "  -t INT    host threads (accepted; alignment, pairing and EM run on the GPU)\n"
    "  -o STR    output prefix [t1k]\n"
    "  -n INT    most alleles a read may be assigned to [2000]\n"
    "  -s FLOAT  minimum alignment similarity [0.8]\n"
    "  --barcode FILE  --alleleWhitelist FILE  --frac FLOAT [0.15]  --cov FLOAT [1.0]  --crossGeneRate FLOAT [0.04]\n"
    "  --relaxIntronAlign  --alleleDigitUnits INT  --alleleDelimiter CHR  --outputReadAssignment  --squaremMinAlpha FLOAT\n";

static void PrintLog(const char *fmt, ...) {
  char msg[2048], stamp[64];
  va_list args;
  va_start(args, fmt);
  vsnprintf(msg, sizeof(msg), fmt, args);
  va_end(args);
  time_t now = time(NULL);
  strftime(stamp, sizeof(stamp), "%c", localtime(&now));
  fprintf(stderr, "[%s] %s\n", stamp, msg);
}

#define T1K_CALL(call)                                                        \
  do {                                                                        \
    int rc_ = (call);                                                         \
    if (rc_ != T1K_OK) {                                                      \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, t1k_last_error()); \
      return EXIT_FAILURE; /* run-t1k dies on a non-zero exit (run-t1k:61-67) */ \
    }                                                                         \
  } while (0)

struct ReadRec {
  char *id, *seq;
  int barcode;
  bool hasN, fragmentAssigned;
};

static ReadRec make_read(const ReadFiles &f, int barcode) {
  ReadRec r;
  r.id = strdup(f.id);
  r.seq = strdup(f.seq);
  r.barcode = barcode;
  r.hasN = strchr(r.seq, 'N') != NULL;
  r.fragmentAssigned = false;
  return r;
}

int main(int argc, char *argv[]) {
  if (argc <= 1) { fprintf(stderr, "%s", usage); return 0; }
  static const char *short_options = "f:a:u:1:2:o:t:n:s:b:";
  static struct option long_options[] = {
      {"frac", required_argument, 0, 10000},          {"cov", required_argument, 0, 10001},
      {"crossGeneRate", required_argument, 0, 10002}, {"barcode", required_argument, 0, 10003},
      {"relaxIntronAlign", no_argument, 0, 10004},    {"alleleDigitUnits", required_argument, 0, 10005},
      {"alleleDelimiter", required_argument, 0, 10006}, {"alleleWhitelist", required_argument, 0, 10007},
      {"outputReadAssignment", no_argument, 0, 10008}, {"squaremMinAlpha", required_argument, 0, 10009},
      {(char *)0, 0, 0, 0}};
  char outputPrefix[1024] = "t1k";
  char buffer[2048];
  Genotyper genotyper(11);
  ReadFiles reads, mateReads, barcodeFile;
  char *refFile = NULL;
  bool hasMate = false, hasBarcode = false, relaxIntronAlign = false, outputReadAssignment = false;
  int maxAssignCnt = 2000, alleleDigitUnits = -1;
  char alleleDelimiter = '\0';
  FILE *fpAbundance = NULL, *fpAlleleWhitelist = NULL;
  double filterFrac = 0.15, filterCov = 1.0, crossGeneRate = 0.04, similarity = 0.8, minSquaremAlpha = 0;
  for (;;) {
    int idx = 0;
    int c = getopt_long(argc, argv, short_options, long_options, &idx);
    if (c == -1) break;
    switch (c) {
      case 'f': refFile = strdup(optarg); break;
      case 'a': fpAbundance = fopen(optarg, "r"); break;
      case 'u': reads.AddReadFile(optarg, false); break;
      case '1': reads.AddReadFile(optarg, true); break;
      case '2': mateReads.AddReadFile(optarg, true); hasMate = true; break;
      case 'o': strcpy(outputPrefix, optarg); break;
      case 't': break;                                  // host threads: nothing on this path is host-parallel
      case 'n': maxAssignCnt = atoi(optarg); break;
      case 's': similarity = atof(optarg); break;
      case 10000: filterFrac = atof(optarg); break;
      case 10001: filterCov = atof(optarg); break;
      case 10002: crossGeneRate = atof(optarg); break;
      case 10003: barcodeFile.AddReadFile(optarg, false); hasBarcode = true; break;
      case 10004: relaxIntronAlign = true; break;
      case 10005: alleleDigitUnits = atoi(optarg); break;
      case 10006: alleleDelimiter = optarg[0]; break;
      case 10007: fpAlleleWhitelist = fopen(optarg, "r"); break;
      case 10008: outputReadAssignment = true; break;
      case 10009: minSquaremAlpha = atof(optarg); genotyper.SetMinSquaremAlpha(minSquaremAlpha); break;
      default: fprintf(stderr, "%s", usage); return EXIT_FAILURE;
    }
  }
  if (refFile == NULL) { fprintf(stderr, "Need to use -f to specify the reference sequences.\n"); return EXIT_FAILURE; }
  genotyper.SetFilterFrac(filterFrac);
  genotyper.SetFilterCov(filterCov);
  genotyper.SetCrossGeneRate(crossGeneRate);
  genotyper.SetAlleleNameStructure(alleleDigitUnits, alleleDelimiter);
  genotyper.InitRefSet(refFile);
  if (fpAlleleWhitelist != NULL) { genotyper.SetAlleleWhitelist(fpAlleleWhitelist); fclose(fpAlleleWhitelist); }
  SeqSet &refSet = genotyper.refSet;
  if (refSet.Size() == 0) { fprintf(stderr, "Need to use -f to specify the reference sequences.\n"); return EXIT_FAILURE; }
  refSet.SetRefSeqSimilarity(similarity);
  refSet.SetRelaxIntronAlign(relaxIntronAlign);
  const int alleleCnt = refSet.Size();

  // ---- reads (Genotyper.cpp:363-442)
  std::vector<ReadRec> reads1, reads2;
  std::map<std::string, int> barcodeStrToInt;
  std::vector<std::string> barcodeIntToStr;
  int maxReadLength = 0;
  while (reads.Next()) {
    int barcode = -1;
    if (hasBarcode) {
      barcodeFile.Next();
      if (!strcmp(barcodeFile.seq, "missing_barcode")) { if (hasMate) mateReads.Next(); continue; }
      std::string s(barcodeFile.seq);
      std::map<std::string, int>::iterator it = barcodeStrToInt.find(s);
      if (it != barcodeStrToInt.end()) barcode = it->second;
      else { barcode = (int)barcodeIntToStr.size(); barcodeStrToInt[s] = barcode; barcodeIntToStr.push_back(s); }
    }
    reads1.push_back(make_read(reads, barcode));
    maxReadLength = std::max(maxReadLength, (int)strlen(reads1.back().seq));
    if (hasMate) {
      mateReads.Next();
      reads2.push_back(make_read(mateReads, barcode));
      maxReadLength = std::max(maxReadLength, (int)strlen(reads2.back().seq));
    }
  }
  genotyper.SetReadLength(maxReadLength);
  const int readCnt = (int)reads1.size();
  genotyper.InitReadAssignments(readCnt, maxAssignCnt);
  PrintLog("Found %d read fragments. Start read assignment.", readCnt);

  // ---- the allele set goes to the device: SeqSet::InputRefSeq's result (SeqSet.hpp:906-982) as plain arrays
  std::string refBases;
  std::vector<int64_t> refOff(alleleCnt + 1, 0);
  std::vector<int32_t> exonPtr(alleleCnt + 1, 0), exonSE;
  for (int i = 0; i < alleleCnt; ++i) {
    refBases.append(refSet.GetSeqConsensus(i), refSet.GetSeqConsensusLen(i));
    refOff[i + 1] = (int64_t)refBases.size();
    const std::vector<struct _pair> &ex = refSet.seqs[i].exons;
    for (size_t e = 0; e < ex.size(); ++e) { exonSE.push_back(ex[e].a); exonSE.push_back(ex[e].b); }
    exonPtr[i + 1] = (int32_t)(exonSE.size() / 2);
  }
  if (exonSE.empty()) exonSE.push_back(0);
  T1KRefDesc desc;
  desc.n_alleles = alleleCnt; desc.bases = refBases.data(); desc.offset = refOff.data();
  desc.exon_ptr = exonPtr.data(); desc.exon_se = exonSE.data();
  desc.similarity = similarity; desc.relax_intron = relaxIntronAlign ? 1 : 0;
  desc.device = getenv("T1K_DEVICE") ? atoi(getenv("T1K_DEVICE")) : 0;
  T1KRef *ref = NULL;
  T1K_CALL(t1k_ref_create(&desc, &ref));

  int alignedFragmentCnt = 0;
  bool emDone = false;
  int emIterCntLib = 0;
  std::vector<double> libRc;
  std::vector<int32_t> libEc, libEcPtr, libEcAlleles, libMissing;
  std::vector<uint8_t> libKept;
  int libEcCnt = 0;
  const bool pipelined = !outputReadAssignment && getenv("T1K_DROPIN_SYNC") == NULL;
  const double tHot0 = (double)clock() / CLOCKS_PER_SEC;
  struct timespec tw0; clock_gettime(CLOCK_MONOTONIC, &tw0);
  if (pipelined) {
    // ---- PHASE A + B + coalescing in ONE call: t1k_genotype runs the chunk pipeline (host de-duplication | device alignment +
    // pairing | host coalescing overlapped) and hands back the coalesced read groups = Genotyper::readAssignments.  The
    // per-fragment rows never leave the library, so this path is taken when --outputReadAssignment does not ask for them.
    const uint32_t stride = (uint32_t)maxReadLength + 1;
    std::vector<char> buf1((size_t)readCnt * stride, 0), buf2(hasMate ? (size_t)readCnt * stride : 0, 0);
    for (int i = 0; i < readCnt; ++i) {
      strcpy(&buf1[(size_t)i * stride], reads1[i].seq);
      if (hasMate) strcpy(&buf2[(size_t)i * stride], reads2[i].seq);
    }
    std::vector<int32_t> seqWeight(alleleCnt), effLen(alleleCnt), alleleMajor(alleleCnt), alleleGene(alleleCnt);
    for (int i = 0; i < alleleCnt; ++i) {
      seqWeight[i] = refSet.GetSeqWeight(i); effLen[i] = refSet.GetSeqEffectiveLen(i);
      alleleMajor[i] = genotyper.alleleInfo[i].majorAlleleIdx; alleleGene[i] = genotyper.alleleInfo[i].geneIdx;
    }
    T1KGroups *grp = NULL;
    T1KGenotypeParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.max_assign = maxAssignCnt; prm.min_squarem_alpha = minSquaremAlpha; prm.filter_frac = filterFrac;
    prm.seq_weight = seqWeight.data(); prm.effective_len = effLen.data();
    prm.allele_major = alleleMajor.data(); prm.allele_gene = alleleGene.data();
    prm.n_major = genotyper.majorAlleleCnt; prm.n_gene = genotyper.geneCnt;
    prm.em_fast_sums = 0; prm.comm = NULL; prm.groups_out = &grp;
    std::vector<double> abundance(alleleCnt), ecAbundance(alleleCnt);
    std::vector<uint8_t> fragAssigned(readCnt > 0 ? readCnt : 1, 0);
    libRc.assign(alleleCnt, 0.0); libEc.assign(alleleCnt, -1);
    libEcPtr.assign(alleleCnt + 1, 0); libEcAlleles.assign(alleleCnt > 0 ? alleleCnt : 1, 0); libMissing.assign(alleleCnt, 0);
    T1KGenotypeResult res;
    memset(&res, 0, sizeof(res));
    res.abundance = abundance.data(); res.ec_abundance = ecAbundance.data(); res.equivalent_class = libEc.data();
    res.missing_coverage = libMissing.data(); res.fragment_assigned = fragAssigned.data(); res.ec_read_count = libRc.data();
    res.ec_allele_ptr = libEcPtr.data(); res.ec_alleles = libEcAlleles.data();
    libKept.assign(alleleCnt > 0 ? alleleCnt : 1, 0); res.allele_kept = libKept.data();
    T1K_CALL(t1k_genotype(ref, buf1.data(), hasMate ? buf2.data() : NULL, stride, (uint32_t)readCnt, &prm, &res));
    int32_t nG = 0; uint64_t nE = 0, nAssigned = 0;
    T1K_CALL(t1k_groups_fetch(grp, &nG, &nE, &nAssigned, NULL, NULL));
    std::vector<int64_t> gptr((size_t)nG + 1);
    std::vector<T1KReadAssignment> gent(nE > 0 ? nE : 1);
    T1K_CALL(t1k_groups_fetch(grp, &nG, &nE, &nAssigned, gptr.data(), gent.data()));
    t1k_groups_destroy(grp);
    static_assert(sizeof(struct _readAssignment) == sizeof(T1KReadAssignment), "T1KReadAssignment mirrors _readAssignment");
    genotyper.readAssignments.resize(nG);
    for (int g = 0; g < nG; ++g) {
      std::vector<struct _readAssignment> &dst = genotyper.readAssignments[g];
      dst.resize((size_t)(gptr[g + 1] - gptr[g]));
      if (!dst.empty()) memcpy(&dst[0], &gent[gptr[g]], dst.size() * sizeof(T1KReadAssignment));
    }
    genotyper.readCnt = nG;
    for (int i = 0; i < readCnt; ++i) reads1[i].fragmentAssigned = fragAssigned[i] != 0;
    alignedFragmentCnt = (int)nAssigned;
    emDone = true; emIterCntLib = res.em_iterations; libEcCnt = res.n_ec;
    if (getenv("T1K_TIMING"))
      fprintf(stderr, "[t1k drop-in] t1k_genotype: dedup %.0f ms (waited %.0f), align %.0f, pair %.0f, coalesce %.0f, em %.0f\n", res.ms_dedup, res.ms_prep_wait,
              res.ms_align, res.ms_pair, res.ms_coalesce, res.ms_em);
  }
  // ---- PHASE A + B in fragment chunks; coalescing stays the reference's (serial, order-sensitive)
  FILE *fpAssign = NULL;
  if (outputReadAssignment) { snprintf(buffer, sizeof(buffer), "%s_assign.tsv", outputPrefix); fpAssign = fopen(buffer, "w"); }
  const int coalesceSize = 500000;          // Genotyper.cpp:523
  const int chunk = 250000;                 // device batch (two per coalescing block)
  std::unordered_map<std::string, uint32_t> uniq;
  std::vector<const char *> uniqSeq;
  std::vector<int32_t> weight;
  std::vector<uint32_t> e1, e2, len;
  std::vector<uint64_t> off;
  std::vector<uint8_t> hasN, assigned;
  std::string bases;
  for (int start = 0; start < readCnt && !pipelined; start += coalesceSize) {
    const int end = std::min(start + coalesceSize, readCnt);       // [start, end)
    for (int c0 = start; c0 < end; c0 += chunk) {
      const int c1 = std::min(c0 + chunk, end), m = c1 - c0;
      uniq.clear(); uniqSeq.clear(); weight.clear();
      e1.assign(m, 0); e2.assign(hasMate ? m : 0, 0); hasN.assign(m, 0); assigned.assign(m, 0);
      for (int i = 0; i < m; ++i) {
        for (int mate = 0; mate < (hasMate ? 2 : 1); ++mate) {
          const ReadRec &r = mate ? reads2[c0 + i] : reads1[c0 + i];
          std::pair<std::unordered_map<std::string, uint32_t>::iterator, bool> ins =
              uniq.insert(std::make_pair(std::string(r.seq), (uint32_t)uniqSeq.size()));
          if (ins.second) { uniqSeq.push_back(r.seq); weight.push_back(0); }
          ++weight[ins.first->second];                              // Genotyper.cpp:149,472: weight = #duplicates
          (mate ? e2 : e1)[i] = ins.first->second;
          if (r.hasN) hasN[i] = 1;
        }
      }
      bases.clear(); off.resize(uniqSeq.size()); len.resize(uniqSeq.size());
      for (size_t k = 0; k < uniqSeq.size(); ++k) {
        off[k] = bases.size(); len[k] = (uint32_t)strlen(uniqSeq[k]);
        bases.append(uniqSeq[k], len[k]);
      }
      T1KAssignment *a = NULL;
      T1K_CALL(t1k_assign_batch(ref, bases.data(), off.data(), len.data(), weight.data(), (uint32_t)uniqSeq.size(), &a));
      uint64_t *rowPtr = NULL;
      T1KReadAssignment *rows = NULL;
      T1K_CALL(t1k_pair_batch(ref, a, e1.data(), hasMate ? e2.data() : NULL, hasN.data(), (uint32_t)m, maxAssignCnt, &rowPtr, &rows,
                              assigned.data()));
      t1k_assignment_destroy(a);
      for (int i = 0; i < m; ++i) {
        std::vector<struct _readAssignment> &dst = genotyper.allReadAssignments[c0 + i];
        dst.resize(rowPtr[i + 1] - rowPtr[i]);
        static_assert(sizeof(struct _readAssignment) == sizeof(T1KReadAssignment), "T1KReadAssignment mirrors _readAssignment");
        if (!dst.empty()) memcpy(&dst[0], rows + rowPtr[i], dst.size() * sizeof(T1KReadAssignment));
        reads1[c0 + i].fragmentAssigned = assigned[i] != 0;
        if (fpAssign)
          for (size_t j = 0; j < dst.size(); ++j)
            fprintf(fpAssign, "%s\t%s\t%d\t%d\n", reads1[c0 + i].id, refSet.GetSeqName(dst[j].alleleIdx), dst[j].start, dst[j].end);
      }
      t1k_free(rowPtr);
      t1k_free(rows);
    }
    alignedFragmentCnt += genotyper.CoalesceReadAssignments(start, end - 1);
  }
  PrintLog("Finish read end assignments.");
  if (fpAssign) fclose(fpAssign);

  // ---- FinalizeReadAssignments (Genotyper.hpp:912-939).  Pipelined path: the library has already built what it builds — the
  // equivalence classes in the reference's own order and every allele's missing coverage — so the driver only fills the
  // reference's structures (T1K_DROPIN_REF_FINALIZE=1 runs the reference's serial code instead: same result, seconds slower).
  if (pipelined && getenv("T1K_DROPIN_REF_FINALIZE") == NULL) {
    for (int g = 0; g < genotyper.readCnt; ++g) {
      const std::vector<struct _readAssignment> &ra = genotyper.readAssignments[g];
      for (size_t j = 0; j < ra.size(); ++j) { struct _pair np; np.a = g; np.b = (int)j; genotyper.readsInAllele[ra[j].alleleIdx].push_back(np); }
    }
    genotyper.equivalentClassToAlleles.clear();
    for (int e = 0; e < libEcCnt; ++e)
      genotyper.equivalentClassToAlleles.push_back(std::vector<int>(libEcAlleles.begin() + libEcPtr[e], libEcAlleles.begin() + libEcPtr[e + 1]));
    for (int i = 0; i < alleleCnt; ++i) { genotyper.alleleInfo[i].equivalentClass = libEc[i]; genotyper.alleleInfo[i].missingCoverage = libMissing[i]; }
    genotyper.RemoveLowMAPQAlleleInEquivalentClass();              // (end of BuildAlleleEquivalentClass, Genotyper.hpp:1136)
  } else {
    std::vector<int32_t> cov(refBases.size());
    T1K_CALL(t1k_coverage_fetch(ref, cov.data()));
    for (int i = 0; i < alleleCnt; ++i) {
      const char *s = refSet.GetSeqConsensus(i);
      const int n = refSet.GetSeqConsensusLen(i);
      for (int j = 0; j < n; ++j)
        if (s[j] != 'N') refSet.seqs[i].posWeight[j].count[(int)nucToNum[s[j] - 'A']] = cov[refOff[i] + j];
    }
    genotyper.FinalizeReadAssignments();
  }
  PrintLog("Finish read fragment assignments. %d read fragments can be assigned (average %.2lf alleles/read).", alignedFragmentCnt,
           genotyper.GetAverageReadAssignmentCnt());

  // ---- PHASE C
  bool sameEc = emDone;
  for (int i = 0; i < alleleCnt && sameEc; ++i) sameEc = libEc[i] == genotyper.alleleInfo[i].equivalentClass;
  if (getenv("T1K_TIMING")) {
    struct timespec tw1; clock_gettime(CLOCK_MONOTONIC, &tw1);
    fprintf(stderr, "[t1k drop-in] hot path + FinalizeReadAssignments: %.0f ms wall (%s path)\n",
            (tw1.tv_sec - tw0.tv_sec) * 1e3 + (tw1.tv_nsec - tw0.tv_nsec) / 1e6, pipelined ? "pipelined" : "synchronous");
    (void)tHot0;
  }
  if (fpAbundance) genotyper.InitAlleleAbundance(fpAbundance);
  else if (sameEc) {
    // the library's EM already ran inside t1k_genotype on the very equivalence classes the reference's FinalizeReadAssignments
    // has just rebuilt (same members, same numbering): its ecReadCount goes straight into the reference's SetAlleleAbundance
    const int E = (int)genotyper.equivalentClassToAlleles.size();
    struct _ecInfo *ecInfo = new struct _ecInfo[E > 0 ? E : 1];
    for (int e = 0; e < E; ++e) {
      const std::vector<int> &members = genotyper.equivalentClassToAlleles[e];
      int length = refSet.GetSeqEffectiveLen(members[0]), missing = genotyper.alleleInfo[members[0]].missingCoverage;
      for (size_t j = 0; j < members.size(); ++j) {
        length = std::min(length, refSet.GetSeqEffectiveLen(members[j]));
        missing = std::min(missing, genotyper.alleleInfo[members[j]].missingCoverage);
      }
      ecInfo[e].length = length; ecInfo[e].missingCoverage = missing;
    }
    genotyper.SetAlleleAbundance(libRc.data(), ecInfo);           // Genotyper.hpp:1316
    delete[] ecInfo;
    PrintLog("Finish allele quantification in %d EM iterations.", emIterCntLib);
  } else {
    // inputs exactly as QuantifyAlleleEquivalentClass assembles them (Genotyper.hpp:1155-1232)
    const int G = genotyper.readCnt, E = (int)genotyper.equivalentClassToAlleles.size();
    std::vector<int64_t> rowPtr(1, 0);
    std::vector<int32_t> col, ecLen(E), ecAllelePtr(1, 0), ecAlleles, alleleMajor(alleleCnt), alleleGene(alleleCnt), stamp(E, -1);
    std::vector<double> count(G), x0(E), x(E), rc(E);
    for (int g = 0; g < G; ++g) {
      const std::vector<struct _readAssignment> &ra = genotyper.readAssignments[g];
      double cnt = ra[0].weight;
      for (size_t j = 1; j < ra.size(); ++j) if (ra[j].weight > cnt) cnt = ra[j].weight;
      count[g] = cnt;
      for (size_t j = 0; j < ra.size(); ++j) {
        const int e = genotyper.alleleInfo[ra[j].alleleIdx].equivalentClass;
        if (stamp[e] != g) { stamp[e] = g; col.push_back(e); }
      }
      rowPtr.push_back((int64_t)col.size());
    }
    struct _ecInfo *ecInfo = new struct _ecInfo[E > 0 ? E : 1];
    for (int e = 0; e < E; ++e) {
      const std::vector<int> &members = genotyper.equivalentClassToAlleles[e];
      int length = refSet.GetSeqEffectiveLen(members[0]), missing = genotyper.alleleInfo[members[0]].missingCoverage;
      double w = 0;
      for (size_t j = 0; j < members.size(); ++j) {
        length = std::min(length, refSet.GetSeqEffectiveLen(members[j]));
        missing = std::min(missing, genotyper.alleleInfo[members[j]].missingCoverage);
        w += refSet.GetSeqWeight(members[j]);
        ecAlleles.push_back(members[j]);
      }
      ecAllelePtr.push_back((int32_t)ecAlleles.size());
      ecInfo[e].length = length; ecInfo[e].missingCoverage = missing;
      ecLen[e] = length; x0[e] = w;
    }
    for (int i = 0; i < alleleCnt; ++i) {
      alleleMajor[i] = genotyper.alleleInfo[i].majorAlleleIdx;
      alleleGene[i] = genotyper.alleleInfo[i].geneIdx;
    }
    int emIterCnt = 0;
    if (E > 0) {
      T1KEmProblem p;
      memset(&p, 0, sizeof(p));
      p.n_groups = G; p.n_ec = E;
      p.row_ptr = rowPtr.data(); p.col = col.data(); p.count = count.data(); p.ec_len = ecLen.data(); p.x0 = x0.data();
      p.min_squarem_alpha = minSquaremAlpha; p.filter_frac = filterFrac;
      p.n_alleles = alleleCnt; p.n_major = genotyper.majorAlleleCnt; p.n_gene = genotyper.geneCnt;
      p.ec_allele_ptr = ecAllelePtr.data(); p.ec_alleles = ecAlleles.data();
      p.allele_major = alleleMajor.data(); p.allele_gene = alleleGene.data();
      p.fast_sums = 0;                               // the reference's summation order: identical doubles
      T1KEmResult r;
      memset(&r, 0, sizeof(r));
      r.x = x.data(); r.ec_read_count = rc.data();
      T1K_CALL(t1k_em_run(&p, &r, desc.device));
      emIterCnt = r.iterations;
    }
    genotyper.SetAlleleAbundance(rc.data(), ecInfo);           // Genotyper.hpp:1316
    delete[] ecInfo;
    PrintLog("Finish allele quantification in %d EM iterations.", emIterCnt);
  }
  t1k_ref_destroy(ref);

  // ---- downstream of the hot path: the reference's own selection and writers (Genotyper.cpp:647-718)
  // RemoveLowLikelihoodAlleleInEquivalentClass (Genotyper.hpp:1371-1460): the pipelined path already has the answer from the
  // library (allele_kept of t1k_genotype, computed over the same read groups, classes and ecAbundance); T1K_DROPIN_REF_FINALIZE=1
  // and the synchronous path run the reference's own code.
  if (sameEc && !fpAbundance && !libKept.empty() && getenv("T1K_DROPIN_REF_FINALIZE") == NULL) {
    for (size_t e = 0; e < genotyper.equivalentClassToAlleles.size(); ++e) {
      std::vector<int> &m = genotyper.equivalentClassToAlleles[e];
      std::vector<int> keptAlleles;
      for (size_t j = 0; j < m.size(); ++j) if (libKept[m[j]]) keptAlleles.push_back(m[j]);
      m = keptAlleles;
    }
  } else
    genotyper.RemoveLowLikelihoodAlleleInEquivalentClass();
  genotyper.SelectAllelesForGenes();
  const int geneCnt = genotyper.GetGeneCnt();
  char *bufferAllele[3];
  for (int j = 0; j < 3; ++j) bufferAllele[j] = new char[20 * (size_t)refSet.Size() + 40];
  snprintf(buffer, sizeof(buffer), "%s_genotype.tsv", outputPrefix);
  FILE *fpOutput = fopen(buffer, "w");
  for (int i = 0; i < geneCnt; ++i) {
    const int called = genotyper.GetAlleleDescription(i, bufferAllele[0], bufferAllele[1], bufferAllele[2]);
    fprintf(fpOutput, "%s\t%d", genotyper.GetGeneName(i), called);
    for (int j = 0; j < 3; ++j) fprintf(fpOutput, "\t%s", bufferAllele[j]);
    fprintf(fpOutput, "\n");
  }
  fclose(fpOutput);
  for (int j = 0; j < 3; ++j) delete[] bufferAllele[j];
  snprintf(buffer, sizeof(buffer), "%s_allele.tsv", outputPrefix);
  genotyper.OutputRepresentativeAlleles(buffer);
  snprintf(buffer, sizeof(buffer), hasMate ? "%s_aligned_1.fa" : "%s_aligned.fa", outputPrefix);
  fpOutput = fopen(buffer, "w");
  for (int i = 0; i < readCnt; ++i)
    if (reads1[i].fragmentAssigned) fprintf(fpOutput, ">%s\n%s\n", reads1[i].id, reads1[i].seq);
  fclose(fpOutput);
  if (hasMate) {
    snprintf(buffer, sizeof(buffer), "%s_aligned_2.fa", outputPrefix);
    fpOutput = fopen(buffer, "w");
    for (int i = 0; i < readCnt; ++i)
      if (reads1[i].fragmentAssigned) fprintf(fpOutput, ">%s\n%s\n", reads2[i].id, reads2[i].seq);
    fclose(fpOutput);
  }
  if (hasBarcode) {
    snprintf(buffer, sizeof(buffer), "%s_aligned_bc.fa", outputPrefix);
    fpOutput = fopen(buffer, "w");
    for (int i = 0; i < readCnt; ++i)
      if (reads1[i].fragmentAssigned) fprintf(fpOutput, ">%s\n%s\n", reads1[i].id, barcodeIntToStr[reads1[i].barcode].c_str());
    fclose(fpOutput);
  }
  PrintLog("Genotyping finishes.");
  return 0;
}
