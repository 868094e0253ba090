// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// Differential-testing harness around the UNMODIFIED reference headers.  It is compiled from
// the reference sources where they lie (-I/root/reference, see oracle/Makefile) into
// oracle/_ref/ref_harness; no reference source is copied into this repository.  The harness
// only calls the reference's own entry points for the hot path and prints what they return:
//
//   align     AlignAlgo::GlobalAlignment            (/root/reference/AlignAlgo.hpp:215-421)
//   assign    SeqSet::AssignRead                    (/root/reference/SeqSet.hpp:2119-2303)
//   alninfo   AssignRead with weight 0 (the analyzer's call, Analyzer.cpp:142,476) followed by
//             SeqSet::AddOverlapAlignmentInfo          (/root/reference/SeqSet.hpp:2657-2680) on every record
//   reads     ReadFiles::AddReadFile / Next (the kseq reader of /root/reference/ReadFiles.hpp:86-204): one sequence per line
//   genotype  the Genotyper.cpp:450-646 flow        (AssignRead -> ReadAssignmentToFragmentAssignment
//             -> SetReadAssignments -> CoalesceReadAssignments -> FinalizeReadAssignments
//             -> QuantifyAlleleEquivalentClass), dumping every boundary the C ABI exposes.
//
// Doubles are printed with %a (exact hex) so that parity can be judged bit-for-bit.
#define private public
#define protected public
#include "Genotyper.hpp"
#undef private
#undef protected

#include <string>
#include <vector>
#include <map>
#include <algorithm>

// Same tables as /root/reference/Genotyper.cpp:37-42 (the headers declare them extern).
char nucToNum[26] = { 0, -1, 1, -1, -1, -1, 2,
	-1, -1, -1, -1, -1, -1, -1,
	-1, -1, -1, -1, -1, 3,
	-1, -1, -1, -1, -1, -1 } ;
char numToNuc[4] = {'A', 'C', 'G', 'T'} ;

static void die(const char *m) { fprintf(stderr, "ref_harness: %s\n", m); exit(2); }

static void print_overlap(FILE *fp, const struct _overlap &o)
{
	fprintf(fp, "%d %d %d %d %d %d %d %d %d %d %a\n", o.seqIdx, o.readStart, o.readEnd, o.seqStart, o.seqEnd,
		o.strand, o.matchCnt, o.relaxedMatchCnt, o.leftClip, o.rightClip, o.similarity);
}

static int mode_align(int argc, char **argv)
{
	// stdin: lines "T P" ("-" = empty string).  stdout: "score ops" with ops as digits 0-3.
	static char t[1 << 16], p[1 << 16], align[1 << 18];
	while (scanf("%65535s %65535s", t, p) == 2)
	{
		int lt = strcmp(t, "-") ? (int)strlen(t) : 0;
		int lp = strcmp(p, "-") ? (int)strlen(p) : 0;
		int score = AlignAlgo::GlobalAlignment(t, lt, p, lp, align);
		printf("%d ", score);
		int k;
		for (k = 0; align[k] != -1; ++k)
			putchar('0' + align[k]);
		if (k == 0) putchar('-');
		putchar('\n');
	}
	return 0;
}

struct Args
{
	const char *ref, *reads1, *reads2, *out;
	double sim; bool relax; int maxAssign; double minAlpha; bool hasMinAlpha; int dumpCov; int ecFilter;
	double filterFrac;
	Args(): ref(NULL), reads1(NULL), reads2(NULL), out(NULL), sim(0.8), relax(false), maxAssign(2000),
		minAlpha(0), hasMinAlpha(false), dumpCov(0), ecFilter(0), filterFrac(0.15) {}
};

static Args parse(int argc, char **argv)
{
	Args a;
	for (int i = 2; i < argc; ++i)
	{
		std::string s(argv[i]);
		if (s == "-f") a.ref = argv[++i];
		else if (s == "-1") a.reads1 = argv[++i];
		else if (s == "-2") a.reads2 = argv[++i];
		else if (s == "-o") a.out = argv[++i];
		else if (s == "-s") a.sim = atof(argv[++i]);
		else if (s == "-n") a.maxAssign = atoi(argv[++i]);
		else if (s == "--relaxIntronAlign") a.relax = true;
		else if (s == "--frac") a.filterFrac = atof(argv[++i]);
		else if (s == "--squaremMinAlpha") { a.minAlpha = atof(argv[++i]); a.hasMinAlpha = true; }
		else if (s == "--cov") a.dumpCov = 1;
		else if (s == "--ecfilter") a.ecFilter = 1;
		else die("unknown argument");
	}
	if (!a.ref) die("need -f");
	return a;
}

// one sequence per line, optional integer weight after whitespace
static void load_lines(const char *fn, std::vector<std::string> &seqs, std::vector<int> &w)
{
	FILE *fp = fopen(fn, "r");
	if (!fp) die("cannot open reads");
	static char buf[1 << 16];
	while (fgets(buf, sizeof(buf), fp))
	{
		char s[1 << 16]; int weight = 1;
		int n = sscanf(buf, "%s %d", s, &weight);
		if (n < 1) continue;
		seqs.push_back(s); w.push_back(n >= 2 ? weight : 1);
	}
	fclose(fp);
}

static void dump_cov(FILE *fp, SeqSet &refSet)
{
	int n = refSet.Size();
	for (int i = 0; i < n; ++i)
	{
		struct _seqWrapper &s = refSet.seqs[i];
		fprintf(fp, "C %d %d", i, s.consensusLen);
		for (int j = 0; j < s.consensusLen; ++j)
		{
			int v = 0;
			if (s.consensus[j] != 'N') v = s.posWeight[j].count[(int)nucToNum[s.consensus[j] - 'A']];
			fprintf(fp, " %d", v);
		}
		fprintf(fp, "\n");
	}
}

static int mode_assign(int argc, char **argv)
{
	Args a = parse(argc, argv);
	Genotyper g(11);
	g.InitRefSet((char *)a.ref);
	SeqSet &refSet = g.refSet;
	refSet.SetRefSeqSimilarity(a.sim);
	refSet.SetRelaxIntronAlign(a.relax);
	std::vector<std::string> seqs; std::vector<int> w;
	load_lines(a.reads1, seqs, w);
	FILE *fp = a.out ? fopen(a.out, "w") : stdout;
	fprintf(fp, "A %d\n", refSet.Size());
	for (size_t i = 0; i < seqs.size(); ++i)
	{
		std::vector<struct _overlap> out;
		int ret = refSet.AssignRead((char *)seqs[i].c_str(), -1, w[i], out);
		fprintf(fp, "R %d %d %d\n", (int)i, ret, (int)out.size());
		for (size_t j = 0; j < out.size(); ++j)
			print_overlap(fp, out[j]);
	}
	if (a.dumpCov) dump_cov(fp, refSet);
	if (a.out) fclose(fp);
	return 0;
}

static int mode_alninfo(int argc, char **argv)
{
	Args a = parse(argc, argv);
	Genotyper g(11);
	g.InitRefSet((char *)a.ref);
	SeqSet &refSet = g.refSet;
	refSet.SetRefSeqSimilarity(a.sim);
	refSet.SetRelaxIntronAlign(a.relax);
	std::vector<std::string> seqs; std::vector<int> w;
	load_lines(a.reads1, seqs, w);
	FILE *fp = a.out ? fopen(a.out, "w") : stdout;
	fprintf(fp, "A %d\n", refSet.Size());
	for (size_t i = 0; i < seqs.size(); ++i)
	{
		std::vector<struct _overlap> out;
		int ret = refSet.AssignRead((char *)seqs[i].c_str(), -1, 0, out);
		fprintf(fp, "R %d %d %d\n", (int)i, ret, (int)out.size());
		for (size_t j = 0; j < out.size(); ++j)
		{
			print_overlap(fp, out[j]);
			refSet.AddOverlapAlignmentInfo((char *)seqs[i].c_str(), out[j]);
			fputs("L ", fp);
			int k;
			for (k = 0; out[j].align[k] != -1; ++k)
				fputc('0' + out[j].align[k], fp);
			if (k == 0) fputc('-', fp);
			fputc('\n', fp);
			delete[] out[j].align;
		}
	}
	if (a.out) fclose(fp);
	return 0;
}

static int mode_reads(int argc, char **argv)
{
	if (argc < 3) die("reads: need a file");
	ReadFiles reads;
	reads.AddReadFile(argv[2], false);
	while (reads.Next())
		printf("%s\n", reads.seq);
	return 0;
}

struct Rd { std::string seq; int mate, idx, info; bool hasN; };
static bool rd_lt(const Rd &a, const Rd &b) { return strcmp(a.seq.c_str(), b.seq.c_str()) < 0; }

static int mode_genotype(int argc, char **argv)
{
	Args a = parse(argc, argv);
	Genotyper g(11);
	g.SetFilterFrac(a.filterFrac);
	if (a.hasMinAlpha) g.SetMinSquaremAlpha(a.minAlpha);
	g.InitRefSet((char *)a.ref);
	SeqSet &refSet = g.refSet;
	refSet.SetRefSeqSimilarity(a.sim);
	refSet.SetRelaxIntronAlign(a.relax);
	std::vector<std::string> s1, s2; std::vector<int> w1, w2;
	load_lines(a.reads1, s1, w1);
	bool hasMate = a.reads2 != NULL;
	if (hasMate) { load_lines(a.reads2, s2, w2); if (s1.size() != s2.size()) die("mate count mismatch"); }
	int readCnt = (int)s1.size();
	int maxLen = 0;
	std::vector<Rd> r1(readCnt), r2(hasMate ? readCnt : 0), all;
	for (int i = 0; i < readCnt; ++i)
	{
		r1[i].seq = s1[i]; r1[i].mate = 0; r1[i].idx = i; r1[i].hasN = s1[i].find('N') != std::string::npos;
		maxLen = std::max(maxLen, (int)s1[i].size());
		all.push_back(r1[i]);
	}
	for (int i = 0; hasMate && i < readCnt; ++i)
	{
		r2[i].seq = s2[i]; r2[i].mate = 1; r2[i].idx = i; r2[i].hasN = s2[i].find('N') != std::string::npos;
		maxLen = std::max(maxLen, (int)s2[i].size());
	}
	all.insert(all.end(), r2.begin(), r2.end());
	g.SetReadLength(maxLen);
	g.InitReadAssignments(readCnt, a.maxAssign);
	std::sort(all.begin(), all.end(), rd_lt);
	int allCnt = (int)all.size();
	std::vector<std::vector<struct _overlap> *> ra(allCnt);
	FILE *fp = a.out ? fopen(a.out, "w") : stdout;
	fprintf(fp, "A %d\n", refSet.Size());
	int uniq = 0;
	for (int i = 0; i < allCnt; )
	{
		int j;
		for (j = i + 1; j < allCnt; ++j)
			if (all[j].seq != all[i].seq) break;
		std::vector<struct _overlap> *v = new std::vector<struct _overlap>;
		refSet.AssignRead((char *)all[i].seq.c_str(), -1, j - i, *v);
		for (int k = i; k < j; ++k) ra[k] = v;
		fprintf(fp, "U %d %s %d %d\n", uniq, all[i].seq.c_str(), j - i, (int)v->size());
		for (size_t k = 0; k < v->size(); ++k) print_overlap(fp, (*v)[k]);
		++uniq;
		i = j;
	}
	for (int i = 0; i < allCnt; ++i)
	{
		if (all[i].mate == 0) r1[all[i].idx].info = i; else r2[all[i].idx].info = i;
	}
	int aligned = 0;
	for (int i = 0; i < readCnt; ++i)
	{
		std::vector<struct _fragmentOverlap> fa;
		bool hasN = r1[i].hasN || (hasMate && r2[i].hasN);
		if (!hasMate) refSet.ReadAssignmentToFragmentAssignment(ra[r1[i].info], NULL, -1, hasN, fa);
		else refSet.ReadAssignmentToFragmentAssignment(ra[r1[i].info], ra[r2[i].info], -1, hasN, fa);
		g.SetReadAssignments(i, fa);
		std::vector<struct _readAssignment> as = g.GetReadAssignments(i);
		fprintf(fp, "F %d %d %d\n", i, (int)fa.size(), (int)as.size());
		for (size_t j = 0; j < fa.size(); ++j)
			fprintf(fp, "f %d %d %d %d %d %a %d\n", fa[j].seqIdx, fa[j].seqStart, fa[j].seqEnd, fa[j].matchCnt,
				fa[j].relaxedMatchCnt, fa[j].similarity, (int)fa[j].hasMatePair);
		for (size_t j = 0; j < as.size(); ++j)
			fprintf(fp, "a %d %d %d %a %a %a\n", as[j].alleleIdx, as[j].start, as[j].end, (double)as[j].weight,
				(double)as[j].qual, (double)as[j].adjustWeight);
	}
	aligned = g.CoalesceReadAssignments(0, readCnt - 1);
	g.FinalizeReadAssignments();
	fprintf(fp, "G %d %d\n", g.readCnt, aligned);
	for (int i = 0; i < g.readCnt; ++i)
	{
		fprintf(fp, "g %d %d", i, (int)g.readAssignments[i].size());
		for (size_t j = 0; j < g.readAssignments[i].size(); ++j)
			fprintf(fp, " %d:%a", g.readAssignments[i][j].alleleIdx, (double)g.readAssignments[i][j].weight);
		fprintf(fp, "\n");
	}
	int ecCnt = (int)g.equivalentClassToAlleles.size();
	fprintf(fp, "E %d\n", ecCnt);
	for (int i = 0; i < ecCnt; ++i)
	{
		fprintf(fp, "e %d %d", i, (int)g.equivalentClassToAlleles[i].size());
		for (size_t j = 0; j < g.equivalentClassToAlleles[i].size(); ++j)
			fprintf(fp, " %d", g.equivalentClassToAlleles[i][j]);
		fprintf(fp, "\n");
	}
	fprintf(fp, "M");
	for (int i = 0; i < g.alleleCnt; ++i) fprintf(fp, " %d", g.alleleInfo[i].missingCoverage);
	fprintf(fp, "\n");
	int iters = ecCnt > 0 ? g.QuantifyAlleleEquivalentClass() : 0;
	fprintf(fp, "Q %d\n", iters);
	for (int i = 0; i < g.alleleCnt; ++i)
		fprintf(fp, "q %d %d %a %a %d %d\n", i, g.alleleInfo[i].equivalentClass, g.alleleInfo[i].abundance,
			g.alleleInfo[i].ecAbundance, refSet.GetSeqEffectiveLen(i), refSet.GetSeqWeight(i));
	if (a.ecFilter)
	{
		// Genotyper.cpp:647: what stays in the classes after the likelihood filter (Genotyper.hpp:1371-1460)
		g.RemoveLowLikelihoodAlleleInEquivalentClass();
		fprintf(fp, "K");
		for (int i = 0; i < ecCnt; ++i)
			for (size_t j = 0; j < g.equivalentClassToAlleles[i].size(); ++j)
				fprintf(fp, " %d", g.equivalentClassToAlleles[i][j]);
		fprintf(fp, "\n");
	}
	if (a.dumpCov) dump_cov(fp, refSet);
	if (a.out) fclose(fp);
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 2) die("usage: ref_harness align|assign|genotype ...");
	std::string m(argv[1]);
	if (m == "align") return mode_align(argc, argv);
	if (m == "assign") return mode_assign(argc, argv);
	if (m == "alninfo") return mode_alninfo(argc, argv);
	if (m == "reads") return mode_reads(argc, argv);
	if (m == "genotype") return mode_genotype(argc, argv);
	die("unknown mode");
	return 2;
}
