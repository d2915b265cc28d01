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
