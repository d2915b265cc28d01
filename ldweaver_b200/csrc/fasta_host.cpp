// Host-side multi-FASTA tokeniser (plain or gz) -> nseq x seq_len byte matrix.
// Stands where the reference uses klib kseq over zlib (src/kseq2.h:167, src/getACGTNsites.cpp:33-45,
// :212-213); one streaming pass instead of the reference's three gzopen passes.  gz inflate is
// sequential CPU work and stays on the host by design.
#include <stdint.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "host_util.h"
#include "../../include/ldw.h"

namespace {

struct Reader {
  gzFile f;
  std::vector<char> buf;
  int len = 0, pos = 0;
  explicit Reader(gzFile f_) : f(f_), buf(1 << 20) {}
  int getc_() {
    if (pos >= len) {
      len = gzread(f, buf.data(), (unsigned)buf.size());
      pos = 0;
      if (len <= 0) return -1;
    }
    return (unsigned char)buf[pos++];
  }
};

}  // namespace

extern "C" int ldw_read_fasta(const char* path, int64_t* nseq_out, int64_t* seq_len_out, uint8_t* aln_out, int64_t aln_cap,
                              char* names_out, int64_t names_cap) {
  if (!path || !nseq_out || !seq_len_out) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: null argument");
  gzFile f = gzopen(path, "rb");
  if (!f) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: can't open %s", path);
  gzbuffer(f, 1 << 20);
  Reader rd(f);
  // second call (aln_out != NULL): *seq_len_out carries the row stride learnt by the query call
  const int64_t stride = aln_out ? *seq_len_out : 0;
  if (aln_out && stride <= 0) {
    gzclose(f);
    return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: pass the queried seq_len in *seq_len_out when aln_out is given");
  }
  int64_t nseq = 0, seq_len = -2, cur_len = 0, names_used = 0;
  bool mismatch = false;
  int c = rd.getc_();
  // skip to first '>'
  while (c != -1 && c != '>') c = rd.getc_();
  while (c == '>') {
    // header: name = up to first whitespace (kseq semantics), rest of the line ignored
    std::string name;
    c = rd.getc_();
    while (c != -1 && c != '\n' && c != ' ' && c != '\t' && c != '\r') { name.push_back((char)c); c = rd.getc_(); }
    while (c != -1 && c != '\n') c = rd.getc_();
    if (names_out) {
      if (names_used + (int64_t)name.size() + 1 > names_cap) {
        gzclose(f);
        return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: names buffer too small");
      }
      memcpy(names_out + names_used, name.c_str(), name.size() + 1);
    }
    names_used += (int64_t)name.size() + 1;
    cur_len = 0;
    c = rd.getc_();
    while (c != -1 && c != '>') {
      if (c != '\n' && c != '\r' && c != ' ' && c != '\t') {  // kseq keeps graphical characters only
        if (aln_out && cur_len < stride) {
          if (nseq * stride + cur_len >= aln_cap) {
            gzclose(f);
            return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: alignment buffer too small");
          }
          aln_out[nseq * stride + cur_len] = (uint8_t)c;
        }
        cur_len++;
      }
      c = rd.getc_();
    }
    if (seq_len == -2) seq_len = cur_len;       // first record defines the length (src/getACGTNsites.cpp:36)
    else if (cur_len != seq_len) mismatch = true;  // :54-56
    nseq++;
  }
  gzclose(f);
  *nseq_out = nseq;
  *seq_len_out = mismatch ? -1 : (seq_len == -2 ? 0 : seq_len);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// ldw_read_fasta_alloc: the same tokeniser in ONE pass over the file.  gz inflate is serial and is the floor (a few
// hundred MB/s), so everything else is kept off its thread: a reader thread inflates into two alternating 16 MB
// buffers while the caller's thread tokenises the previous one, and sequence lines are moved with memchr / memcpy
// instead of a per-character loop.  All sequence bytes are appended to one growing buffer (realloc: mremap for blocks
// of this size, no copy); when every record has the length of the first, that buffer IS the nseq x seq_len matrix.
// ---------------------------------------------------------------------------------------------------------------------
#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

struct ChunkPipe {  // two buffers handed back and forth between the inflating thread and the tokeniser
  static constexpr size_t kCap = size_t(16) << 20;
  std::vector<char> buf[2];
  int len[2] = {0, 0};          // bytes in the buffer; -1 = end of file, -2 = read error
  bool full[2] = {false, false};
  bool stop = false;            // tokeniser gave up: the reader must not block
  std::mutex mu;
  std::condition_variable cv;
};

inline bool is_fasta_space(unsigned char c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t'; }

struct Grow {  // byte buffer grown by realloc
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

}  // namespace

extern "C" void ldw_buffer_free(void* p) { free(p); }

extern "C" int ldw_read_fasta_alloc(const char* path, int64_t* nseq_out, int64_t* seq_len_out, uint8_t** aln_out, char** names_out,
                                    int64_t* names_len_out) {
  return ldw::guarded("ldw_read_fasta_alloc", [&]() -> int {
  if (!path || !nseq_out || !seq_len_out || !aln_out || !names_out || !names_len_out)
    return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta_alloc: null argument");
  *nseq_out = 0; *seq_len_out = 0; *aln_out = nullptr; *names_out = nullptr; *names_len_out = 0;
  gzFile f = gzopen(path, "rb");
  if (!f) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta: can't open %s", path);
  gzbuffer(f, 1 << 20);
  ChunkPipe P;
  P.buf[0].resize(ChunkPipe::kCap);
  P.buf[1].resize(ChunkPipe::kCap);
  std::thread reader([&]() {
    for (int k = 0;; k ^= 1) {
      {
        std::unique_lock<std::mutex> lk(P.mu);
        P.cv.wait(lk, [&] { return !P.full[k] || P.stop; });
        if (P.stop) return;
      }
      int got = gzread(f, P.buf[k].data(), (unsigned)ChunkPipe::kCap);
      {
        std::lock_guard<std::mutex> lk(P.mu);
        P.len[k] = got > 0 ? got : (got == 0 ? -1 : -2);
        P.full[k] = true;
      }
      P.cv.notify_all();
      if (got <= 0) return;
    }
  });
  auto finish_reader = [&]() {
    { std::lock_guard<std::mutex> lk(P.mu); P.stop = true; }
    P.cv.notify_all();
    reader.join();
    gzclose(f);
  };

  Grow seq;                       // all sequence bytes, record after record
  std::string names;              // NUL-terminated names back to back
  int64_t nseq = 0, seq_len = -2, cur_len = 0;
  bool mismatch = false, oom = false, read_error = false;
  enum { BEFORE_FIRST, NAME, HEADER_REST, SEQ } st = BEFORE_FIRST;
  auto end_record = [&]() {
    if (seq_len == -2) seq_len = cur_len;          // first record defines the length (src/getACGTNsites.cpp:36)
    else if (cur_len != seq_len) mismatch = true;  // :54-56
    nseq++;
  };
  for (int k = 0;; k ^= 1) {
    int len;
    {
      std::unique_lock<std::mutex> lk(P.mu);
      P.cv.wait(lk, [&] { return P.full[k]; });
      len = P.len[k];
    }
    if (len < 0) { read_error = (len == -2); break; }
    const char* p = P.buf[k].data();
    const char* const end = p + len;
    while (p < end && !oom) {
      switch (st) {
        case BEFORE_FIRST: {
          const char* q = (const char*)memchr(p, '>', (size_t)(end - p));
          if (!q) { p = end; break; }
          p = q + 1; st = NAME; cur_len = 0;
          break;
        }
        case NAME: {  // name = up to the first whitespace (kseq semantics)
          const char* q = p;
          while (q < end && !is_fasta_space((unsigned char)*q)) q++;
          names.append(p, (size_t)(q - p));
          p = q;
          if (q < end) { names.push_back('\0'); st = (*q == '\n') ? SEQ : HEADER_REST; if (*q == '\n') p++; }
          break;
        }
        case HEADER_REST: {  // rest of the header line is ignored
          const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
          if (!q) { p = end; break; }
          p = q + 1; st = SEQ;
          break;
        }
        case SEQ: {  // up to the next '>' (wherever it stands, as the character loop of ldw_read_fasta): whitespace dropped
          const char* gt = (const char*)memchr(p, '>', (size_t)(end - p));
          const char* const stop = gt ? gt : end;
          if (!seq.reserve((size_t)(stop - p))) { oom = true; break; }
          while (p < stop) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(stop - p));
            const char* const le = nl ? nl : stop;
            size_t m = (size_t)(le - p);
            if (m && p[m - 1] == '\r') m--;  // CRLF
            if (m && (memchr(p, ' ', m) || memchr(p, '\t', m) || memchr(p, '\r', m))) {
              for (size_t i = 0; i < m; i++) if (!is_fasta_space((unsigned char)p[i])) { seq.p[seq.n++] = (uint8_t)p[i]; cur_len++; }
            } else {
              memcpy(seq.p + seq.n, p, m);
              seq.n += m; cur_len += (int64_t)m;
            }
            p = nl ? nl + 1 : stop;
          }
          if (gt) { end_record(); p = gt + 1; st = NAME; cur_len = 0; }
          break;
        }
      }
    }
    {
      std::lock_guard<std::mutex> lk(P.mu);
      P.full[k] = false;
    }
    P.cv.notify_all();
    if (oom) break;
  }
  finish_reader();
  if (oom) return ldw::set_error(LDW_ERR_NOMEM, "ldw_read_fasta_alloc: out of host memory after %lld sequence bytes", (long long)seq.n);
  if (read_error) return ldw::set_error(LDW_ERR_ARG, "ldw_read_fasta_alloc: error while reading %s", path);
  if (st == NAME) { names.push_back('\0'); st = SEQ; }  // file ends inside a header line: a record without sequence
  if (st == SEQ || st == HEADER_REST) end_record();
  *nseq_out = nseq;
  *seq_len_out = mismatch ? -1 : (seq_len == -2 ? 0 : seq_len);
  *names_len_out = (int64_t)names.size();
  char* nm = (char*)malloc(names.size() + 1);
  if (!nm) return ldw::set_error(LDW_ERR_NOMEM, "ldw_read_fasta_alloc: out of host memory");
  memcpy(nm, names.data(), names.size());
  nm[names.size()] = 0;
  *names_out = nm;
  if (!mismatch && seq.n) { *aln_out = seq.p; seq.p = nullptr; }  // ownership moves to the caller (ldw_buffer_free)
  return 0;
  });
}
