// SURVEY.md §8f N2, first half: reads into memory the way Genotyper.cpp:363-454 gets them from ReadFiles (ReadFiles.hpp:155-204,
// a kseq stream over gzopen): FASTA or FASTQ, plain or gzip, sequences exactly as they stand in the file.  The record grammar is
// kseq's: a header starts at '>' or '@', its name ends at the first white space, sequence lines are joined until a line that
// starts with '>', '@' or '+'; after a '+' line quality lines are consumed until they are as long as the sequence (so a quality
// string may start with '@'); a trailing '\r' of a line is dropped; a record whose quality is short ends the file
// (ReadFiles::Next treats kseq's -2 like end of file).  Plain C++ + zlib, no CUDA.
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>

#include <memory>
#include <string>
#include <vector>

namespace t1k {

class SeqFileReader {
 public:
  explicit SeqFileReader(const char *path) : f_(gzopen(path, "r")), buf_(1 << 20) { if (f_) gzbuffer(f_, 1 << 20); }
  ~SeqFileReader() { if (f_) gzclose(f_); }
  bool ok() const { return f_ != nullptr; }
  // next record's sequence appended to `out` (not cleared); returns its length, -1 at end of file, -2 on a truncated quality string
  int64_t next(std::vector<char> &out) {
    int c;
    if (last_ == 0) {
      while ((c = getc()) != -1 && c != '>' && c != '@') {}
      if (c == -1) return -1;
      last_ = c;
    }
    // name up to the first white space, then the rest of the header line
    bool any = false;
    for (;;) {
      c = getc();
      if (c == -1) { if (!any) return -1; break; }
      if (c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r') break;
      any = true;
    }
    if (c != '\n' && c != -1) while ((c = getc()) != -1 && c != '\n') {}
    const size_t start = out.size();
    while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@') {
      if (c == '\n') continue;
      out.push_back((char)c);
      rest_of_line(out, start);
    }
    if (c == '>' || c == '@') last_ = c;
    const int64_t len = (int64_t)(out.size() - start);
    if (c != '+') { if (c == -1) last_ = -1; return len; }
    while ((c = getc()) != -1 && c != '\n') {}
    if (c == -1) return -2;
    qual_.clear();
    while (!eof_line_ && (int64_t)qual_.size() < len) { if (!rest_of_line(qual_, 0)) break; }
    last_ = 0;
    if ((int64_t)qual_.size() != len) return -2;
    return len;
  }

 private:
  gzFile f_;
  std::vector<unsigned char> buf_;
  int begin_ = 0, end_ = 0;
  bool eof_ = false, eof_line_ = false;
  int last_ = 0;
  std::vector<char> qual_;
  int getc() {
    if (last_ == -1) return -1;
    if (begin_ >= end_) {
      if (eof_ || !f_) return -1;
      begin_ = 0;
      end_ = gzread(f_, buf_.data(), (unsigned)buf_.size());
      if (end_ <= 0) { eof_ = true; end_ = 0; return -1; }
    }
    return buf_[begin_++];
  }
  // appends the rest of the current line (without '\n', without a trailing '\r' when the string is longer than one
  // character); false when nothing could be read at all (end of file)
  bool rest_of_line(std::vector<char> &s, size_t base) {
    bool got = false;
    for (;;) {
      if (begin_ >= end_) {
        if (eof_ || !f_) { eof_line_ = true; break; }
        begin_ = 0;
        end_ = gzread(f_, buf_.data(), (unsigned)buf_.size());
        if (end_ <= 0) { eof_ = true; end_ = 0; eof_line_ = true; break; }
      }
      const unsigned char *p = (const unsigned char *)memchr(buf_.data() + begin_, '\n', (size_t)(end_ - begin_));
      const int stop = p ? (int)(p - buf_.data()) : end_;
      s.insert(s.end(), buf_.data() + begin_, buf_.data() + stop);
      got = got || stop > begin_ || p;
      begin_ = p ? stop + 1 : stop;
      if (p) break;
    }
    if (s.size() - base > 1 && s.back() == '\r') s.pop_back();
    return got;
  }
};

struct LoadedReads {
  std::vector<char> bases[2];          // concatenated sequences of mate 1 / mate 2
  std::vector<uint64_t> off[2];        // [n + 1]
  uint32_t maxLen = 0;
};

// 0 ok; 1 cannot open; 2 the mate file holds fewer records; 3 a read is longer than maxAllowed
inline int load_reads(const char *path1, const char *path2, uint32_t maxAllowed, LoadedReads &R, std::string &err) {
  SeqFileReader a(path1);
  if (!a.ok()) { err = std::string("cannot open ") + path1; return 1; }
  std::unique_ptr<SeqFileReader> b;
  if (path2) { b.reset(new SeqFileReader(path2)); if (!b->ok()) { err = std::string("cannot open ") + path2; return 1; } }
  R.off[0].assign(1, 0); R.off[1].assign(1, 0);
  for (;;) {
    const int64_t n = a.next(R.bases[0]);
    if (n < 0) break;
    R.off[0].push_back(R.bases[0].size());
    if ((uint64_t)n > R.maxLen) R.maxLen = (uint32_t)n;
    if (b) {
      const int64_t m = b->next(R.bases[1]);
      if (m < 0) { err = "the mate file holds fewer records than the first file"; return 2; }
      R.off[1].push_back(R.bases[1].size());
      if ((uint64_t)m > R.maxLen) R.maxLen = (uint32_t)m;
    }
    if (R.maxLen > maxAllowed) { err = "read longer than T1K_MAX_READ_LEN"; return 3; }
  }
  return 0;
}

}  // namespace t1k
