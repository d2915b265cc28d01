// Host-side multi-FASTA tokeniser (plain or gz) -> nseq x seq_len byte matrix, in ONE streaming pass.
//
// Stands where the reference runs klib's kseq over zlib three times per file (src/kseq2.h:167-207 `kseq_read`, called
// from src/getACGTNsites.cpp:33-45, :50, :222).  gz inflate is sequential CPU work and stays on the host by design
// (SURVEY.md section 2a).  The record grammar is kseq_read's, byte for byte, because seq.length (= snp.dat$g) and every
// POS depend on it; tests/test_fasta_ref_cpu.py checks this file against the reference's compiled reader
// (oracle/_ref) on CRLF, blank-line, '@' / '+' and random byte streams.  What that grammar is:
//   * before the first record, and after a record that had a '+' section, bytes are skipped up to the next '>' or
//     '@' wherever it stands (kseq2.h:171-174);
//   * the name runs to the first isspace() byte; unless that byte is '\n' the rest of the line is a comment (:177-178);
//   * then line by line: a line whose first byte is '>' or '@' starts the next record, one that starts with '+' opens
//     a FASTQ quality section, any other first byte -- a '\n' of an empty line, '\r', blank -- is APPENDED and so is
//     everything up to, not including, the next '\n' (:183-186).  Hence the '\r' of CRLF files and inner blanks are
//     sequence bytes, and an empty line glues the whole following line (a header included) onto the sequence;
//   * after '+': the rest of that line is dropped, then whole lines are appended to the quality string until it is at
//     least as long as the sequence; a length mismatch or a missing quality section ends the READING (return -2: the
//     reference's `while ((l = kseq_read(seq)) >= 0)` loops stop there, records so far stand) (:196-205);
//   * callers take strlen() of the sequence and the C string of the name (getACGTNsites.cpp:36,51-52), so both stop
//     at an embedded NUL byte.
// One stream-buffer artefact is reproduced too: a header character that is the very last byte of the stream yields no
// record (ks_getuntil sees EOF, :176), unless the stream length is a multiple of kseq's 16384-byte buffer, in which
// case EOF is not yet known and an empty record is returned (:87-98).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_util.h"
#include "../../include/ldw.h"

namespace {

struct Grow {  // byte buffer grown by realloc (mremap for blocks of this size: no copy)
  uint8_t* p = nullptr;
  size_t n = 0, cap = 0;
  bool reserve(size_t extra) {
    if (n + extra <= cap) return true;
    size_t want = std::max(n + extra, cap + cap / 2 + (size_t(1) << 20));
    void* q = realloc(p, want);
    if (!q) return false;
    p = (uint8_t*)q;
    cap = want;
    return true;
  }
  ~Grow() { free(p); }
};

// C-locale isspace(), which is what ks_getuntil(KS_SEP_SPACE) tests (kseq2.h:104)
inline bool c_isspace(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// Push tokeniser: feed() takes the stream in arbitrary pieces, finish() sees the end of the stream.
struct KseqMachine {
  enum State { SKIP_TO_HEADER, NAME, COMMENT, LINE_START, LINE_REST, PLUS_LINE, QUAL, STOPPED } st = SKIP_TO_HEADER;
  Grow seq;                  // sequence bytes of all finished records back to back + the one being read
  std::string names;         // NUL-terminated names of the finished records back to back
  std::string name;          // name being read
  size_t rec_begin = 0;      // where the current record's sequence starts in `seq`
  size_t qual_len = 0;       // bytes of the current quality string
  bool name_started = false; // NAME state has consumed at least one byte of this stream position
  int64_t nseq = 0, seq_len = -2;
  bool mismatch = false, oom = false;

  void end_record() {
    size_t len = seq.n - rec_begin;
    if (len) {  // callers use strlen(seq->seq.s): the record ends at an embedded NUL
      const void* z = memchr(seq.p + rec_begin, 0, len);
      if (z) { len = (size_t)((const uint8_t*)z - (seq.p + rec_begin)); seq.n = rec_begin + len; }
    }
    names.append(name.c_str());  // C string: stops at an embedded NUL as seq->name.s does
    names.push_back('\0');
    if (seq_len == -2) seq_len = (int64_t)len;            // first record defines the length (getACGTNsites.cpp:36)
    else if ((int64_t)len != seq_len) mismatch = true;    // :54-56
    nseq++;
    rec_begin = seq.n;
  }
  void drop_record() { seq.n = rec_begin; }
  void begin_record() { name.clear(); name_started = false; qual_len = 0; st = NAME; }

  void feed(const char* p, size_t n) {
    const char* const end = p + n;
    while (p < end && !oom && st != STOPPED) {
      switch (st) {
        case SKIP_TO_HEADER: {
          const char* a = (const char*)memchr(p, '>', (size_t)(end - p));
          const char* b = (const char*)memchr(p, '@', (size_t)(a ? a - p : end - p));
          const char* q = b ? b : a;
          if (!q) { p = end; break; }
          p = q + 1;
          begin_record();
          break;
        }
        case NAME: {
          name_started = true;
          const char* q = p;
          while (q < end && !c_isspace((unsigned char)*q)) q++;
          name.append(p, (size_t)(q - p));
          p = q;
          if (q < end) { st = (*q == '\n') ? LINE_START : COMMENT; p++; }
          break;
        }
        case COMMENT: {
          const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
          if (!q) { p = end; break; }
          p = q + 1;
          st = LINE_START;
          break;
        }
        case LINE_START: {
          const char c = *p++;
          if (c == '>' || c == '@') { end_record(); begin_record(); break; }
          if (c == '+') { st = PLUS_LINE; break; }
          if (!seq.reserve(1)) { oom = true; break; }
          seq.p[seq.n++] = (uint8_t)c;  // whatever it is, '\n' of an empty line included
          st = LINE_REST;
          break;
        }
        case LINE_REST: {
          const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
          const size_t m = (size_t)((q ? q : end) - p);
          if (!seq.reserve(m)) { oom = true; break; }
          memcpy(seq.p + seq.n, p, m);
          seq.n += m;
          p += m;
          if (q) { p++; st = LINE_START; }
          break;
        }
        case PLUS_LINE: {
          const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
          if (!q) { p = end; break; }
          p = q + 1;
          st = QUAL;
          break;
        }
        case QUAL: {  // whole lines until the quality string is at least as long as the sequence (at least one line)
          const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
          qual_len += (size_t)((q ? q : end) - p);
          if (!q) { p = end; break; }
          p = q + 1;
          if (qual_len >= seq.n - rec_begin) {
            if (qual_len == seq.n - rec_begin) { end_record(); st = SKIP_TO_HEADER; }
            else { drop_record(); st = STOPPED; }  // return -2
          }
          break;
        }
        case STOPPED: break;
      }
    }
  }

  void finish(uint64_t stream_bytes) {
    switch (st) {
      case NAME:
        // ks_getuntil at EOF returns -1 (no record) -- unless EOF is not yet known, which is the case exactly when the
        // stream length is a multiple of kseq's 16384-byte buffer (see the header comment)
        if (name_started || stream_bytes % 16384 == 0) end_record();
        break;
      case COMMENT: case LINE_START: case LINE_REST: end_record(); break;
      case QUAL:
        if (qual_len == seq.n - rec_begin) end_record(); else drop_record();
        break;
      case PLUS_LINE: drop_record(); break;  // return -2: no quality string
      default: break;
    }
    st = STOPPED;
  }
};

struct ChunkPipe {  // two buffers handed back and forth between the inflating thread and the tokeniser
  std::vector<char> buf[2];
  int len[2] = {0, 0};          // bytes in the buffer; -1 = end of file, -2 = read error
  bool full[2] = {false, false};
  bool stop = false;            // tokeniser gave up: the reader must not block
  std::mutex mu;
  std::condition_variable cv;
};

// inflate on a reader thread into two alternating buffers while the caller's thread tokenises the previous one
int read_all(const char* path, KseqMachine& M) {
  gzFile f = gzopen(path, "rb");
  if (!f) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: can't open %s", path);
  gzbuffer(f, 1 << 20);
  size_t cap = size_t(16) << 20;
  if (const char* e = getenv("LDW_FASTA_CHUNK")) {  // test hook: tiny hand-over buffers put every state on a boundary
    long v = atol(e);
    if (v >= 1) cap = (size_t)v;
  }
  ChunkPipe P;
  P.buf[0].resize(cap);
  P.buf[1].resize(cap);
  std::thread reader([&]() {
    for (int k = 0;; k ^= 1) {
      {
        std::unique_lock<std::mutex> lk(P.mu);
        P.cv.wait(lk, [&] { return !P.full[k] || P.stop; });
        if (P.stop) return;
      }
      int got = gzread(f, P.buf[k].data(), (unsigned)cap);
      {
        std::lock_guard<std::mutex> lk(P.mu);
        P.len[k] = got > 0 ? got : (got == 0 ? -1 : -2);
        P.full[k] = true;
      }
      P.cv.notify_all();
      if (got <= 0) return;
    }
  });
  bool read_error = false;
  uint64_t total = 0;
  for (int k = 0;; k ^= 1) {
    int len;
    {
      std::unique_lock<std::mutex> lk(P.mu);
      P.cv.wait(lk, [&] { return P.full[k]; });
      len = P.len[k];
    }
    if (len < 0) { read_error = (len == -2); break; }
    total += (uint64_t)len;
    M.feed(P.buf[k].data(), (size_t)len);
    {
      std::lock_guard<std::mutex> lk(P.mu);
      P.full[k] = false;
    }
    P.cv.notify_all();
    if (M.oom || M.st == KseqMachine::STOPPED) break;
  }
  { std::lock_guard<std::mutex> lk(P.mu); P.stop = true; }
  P.cv.notify_all();
  reader.join();
  gzclose(f);
  if (M.oom) return ldw::set_error(LDW_ERR_NOMEM, "ldw_read_fasta: out of host memory after %lld sequence bytes", (long long)M.seq.n);
  if (read_error) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: error while reading %s", path);
  M.finish(total);
  return 0;
}

}  // namespace

extern "C" void ldw_buffer_free(void* p) { free(p); }

extern "C" int ldw_read_fasta_alloc(const char* path, int64_t* nseq_out, int64_t* seq_len_out, uint8_t** aln_out, char** names_out,
                                    int64_t* names_len_out) {
  return ldw::guarded("ldw_read_fasta_alloc", [&]() -> int {
    if (!path || !nseq_out || !seq_len_out || !aln_out || !names_out || !names_len_out)
      return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta_alloc: null argument");
    *nseq_out = 0; *seq_len_out = 0; *aln_out = nullptr; *names_out = nullptr; *names_len_out = 0;
    KseqMachine M;
    if (int rc = read_all(path, M)) return rc;
    char* nm = (char*)malloc(M.names.size() + 1);
    if (!nm) return ldw::set_error(LDW_ERR_NOMEM, "ldw_read_fasta_alloc: out of host memory");
    memcpy(nm, M.names.data(), M.names.size());
    nm[M.names.size()] = 0;
    *names_out = nm;
    *names_len_out = (int64_t)M.names.size();
    *nseq_out = M.nseq;
    *seq_len_out = M.mismatch ? -1 : (M.seq_len == -2 ? 0 : M.seq_len);
    // every record has the first one's length: the byte buffer IS the nseq x seq_len matrix; ownership moves to the caller
    if (!M.mismatch && M.seq.n) { *aln_out = M.seq.p; M.seq.p = nullptr; }
    return 0;
  });
}

// The two-call form (query, then fill caller memory) over the same tokeniser.
extern "C" int ldw_read_fasta(const char* path, int64_t* nseq_out, int64_t* seq_len_out, uint8_t* aln_out, int64_t aln_cap,
                              char* names_out, int64_t names_cap) {
  return ldw::guarded("ldw_read_fasta", [&]() -> int {
    if (!path || !nseq_out || !seq_len_out) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: null argument");
    const int64_t stride = aln_out ? *seq_len_out : 0;
    if (aln_out && stride <= 0)
      return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: pass the queried seq_len in *seq_len_out when aln_out is given");
    KseqMachine M;
    if (int rc = read_all(path, M)) return rc;
    const int64_t seq_len = M.mismatch ? -1 : (M.seq_len == -2 ? 0 : M.seq_len);
    if (names_out) {
      if ((int64_t)M.names.size() > names_cap) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: names buffer too small");
      memcpy(names_out, M.names.data(), M.names.size());
    }
    if (aln_out) {
      if (seq_len != stride) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: the file no longer has the queried seq_len");
      if ((int64_t)M.seq.n > aln_cap) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: alignment buffer too small");
      memcpy(aln_out, M.seq.p, M.seq.n);
    }
    *nseq_out = M.nseq;
    *seq_len_out = seq_len;
    return 0;
  });
}
