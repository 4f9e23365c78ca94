// dbfile.cu — the packed on-disk target database: FASTA -> digital residues, once, with all host cores;
// afterwards a search maps the file and streams it to the GPU (pipeline.cu) instead of re-parsing text for every query.
//
// What it replaces on the search path: esl_sqfile_OpenDigital + esl_sqio_ReadBlock (HMMER/easel/esl_sqio_ascii.c:1466-1500,
// one reader thread, ~200 M residues/s: the end-to-end ceiling of multi-threaded hmmsearch).  Precedents in
// the reference for a pre-digitised database: src/cachedb.h (P7_SEQCACHE), src/makehmmerdb.c.
//
// Digitisation is easel's, restated: the amino alphabet "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~" with '_' and '.' as gaps and lower
// case accepted (esl_alphabet.c:269-296), the FASTA input map of esl_sqio_ascii.c:2952-2973 (space, tab, CR ignored, LF
// ends a line, '>' ANYWHERE outside a header line ends the record, everything else illegal), the header rule of
// header_fasta (:2997-3056: name = first word after '>', description = rest of the line).  Host code only; no kernels.
#include "p7gpu_internal.h"
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

namespace {

constexpr uint8_t C_ILLEGAL = 254, C_IGNORED = 253, C_EOL = 252, C_EOD = 251, C_SENT = 255;
constexpr char MAGIC[8] = {'P', '7', 'G', 'P', 'U', 'D', 'B', '1'};

struct Header {                    // 128 bytes, little-endian
  char     magic[8];
  uint32_t version, Kp;
  uint64_t nseq, nbytes, nres;
  int32_t  max_length, pad0;
  uint64_t off_residues, off_offsets, off_lengths, off_nameidx, off_names, names_bytes;
  uint64_t reserved[4];
};
static_assert(sizeof(Header) == 128, "header layout");

struct InMap {
  uint8_t m[256];
  InMap() {
    for (int x = 0; x < 256; x++) m[x] = C_ILLEGAL;
    const char *sym = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~";          // esl_alphabet.c:275
    for (int x = 0; sym[x]; x++) {
      m[(unsigned char) sym[x]] = (uint8_t) x;
      if (sym[x] >= 'A' && sym[x] <= 'Z') m[(unsigned char) (sym[x] - 'A' + 'a')] = (uint8_t) x;   // SetCaseInsensitive, :282
    }
    m[(unsigned char) '_'] = m[(unsigned char) '-'];            // :280-281
    m[(unsigned char) '.'] = m[(unsigned char) '-'];
    m[(unsigned char) ' '] = C_IGNORED; m[(unsigned char) '\t'] = C_IGNORED; m[(unsigned char) '\r'] = C_IGNORED;   // esl_sqio_ascii.c:2966-2970
    m[(unsigned char) '\n'] = C_EOL; m[(unsigned char) '>'] = C_EOD;
  }
};
const InMap g_inmap;

struct Chunk {                     // what one thread produced for its byte range of the text
  std::vector<uint8_t> res;        // residues of its records, each followed by one sentinel
  std::vector<int32_t> len;
  std::vector<char>    names;      // name\0desc\0 per record
  std::vector<uint64_t> nameoff;
  int64_t err_pos = -1; int err_char = 0; int err_kind = 0;   // first error in this chunk (byte position in the text)
};

// true if byte position p lies inside a header line (between a record's '>' and the end of that line)
bool in_header(const char *t, int64_t p)
{
  int64_t ls = p;
  while (ls > 0 && t[ls - 1] != '\n' && t[ls - 1] != '\r') ls--;
  for (int64_t q = ls; q < p; q++) if (t[q] == '>') return true;
  return false;
}

// first record start ('>' outside a header line) at or after p, or n
int64_t next_record(const char *t, int64_t n, int64_t p)
{
  bool hdr = in_header(t, p);
  for (; p < n; p++) {
    const char c = t[p];
    if (hdr) { if (c == '\n' || c == '\r') hdr = false; }
    else if (c == '>') return p;
  }
  return n;
}

void parse_range(const char *t, int64_t n, int64_t lo, int64_t hi, int align, Chunk *out)
{
  int64_t p = next_record(t, n, lo);
  while (p < hi && p < n) {
    // ---- header line (header_fasta, esl_sqio_ascii.c:3014-3045) ----
    int64_t q = p + 1;
    while (q < n && (t[q] == ' ' || t[q] == '\t')) q++;
    const int64_t name0 = q;
    while (q < n && !isspace((unsigned char) t[q])) q++;
    if (q == name0) { out->err_pos = p; out->err_kind = 2; return; }           // "no FASTA name found"
    out->nameoff.push_back(out->names.size());
    out->names.insert(out->names.end(), t + name0, t + q); out->names.push_back('\0');
    while (q < n && (t[q] == ' ' || t[q] == '\t')) q++;
    const int64_t d0 = q;
    while (q < n && t[q] != '\n' && t[q] != '\r') q++;
    out->names.insert(out->names.end(), t + d0, t + q); out->names.push_back('\0');
    while (q < n && (t[q] == '\n' || t[q] == '\r')) q++;
    if (q >= n) { out->err_pos = p; out->err_kind = 4; return; }               // header line runs into the end of the file, :3047
    // ---- sequence data up to the next '>' or the end of the text ----
    const size_t r0 = out->res.size();
    for (; q < n; q++) {
      const unsigned char c = (unsigned char) t[q];
      const uint8_t x = g_inmap.m[c];
      if (x < 29) out->res.push_back(x);
      else if (x == C_EOD) break;
      else if (x == C_ILLEGAL) { out->err_pos = q; out->err_char = c; out->err_kind = (c >= 128) ? 3 : 1; return; }
    }
    if (out->res.size() - r0 > (size_t) 0x7fffffff) { out->err_pos = p; out->err_kind = 5; return; }   // ESL_SQ.n is 64-bit; the GPU path is not
    out->len.push_back((int32_t) (out->res.size() - r0));
    out->res.push_back(C_SENT);
    while (out->res.size() % (size_t) align) out->res.push_back(C_SENT);      // the next target starts on an `align` boundary
    p = q;
  }
}

int64_t line_of(const char *t, int64_t pos) { int64_t l = 1; for (int64_t i = 0; i < pos; i++) l += (t[i] == '\n'); return l; }

int write_all(int fd, const void *buf, size_t n)
{
  const char *p = (const char *) buf;
  while (n) { ssize_t w = write(fd, p, std::min<size_t>(n, (size_t) 1 << 30)); if (w <= 0) return -1; p += w; n -= (size_t) w; }
  return 0;
}
int pad_to(int fd, uint64_t *pos, uint64_t align)
{
  static const char z[4096] = {0};
  const uint64_t r = (*pos) % align;
  if (r) { if (write_all(fd, z, (size_t) (align - r))) return -1; *pos += align - r; }
  return 0;
}

}  // namespace

struct p7gpu_dbfile {
  void *map = nullptr; size_t map_bytes = 0; int fd = -1;
  const Header *h = nullptr;
  bool pinned = false;
  uint8_t *p_res = nullptr; int64_t *p_off = nullptr; int32_t *p_len = nullptr;   // page-locked copies (p7gpu_dbfile_pin)
  uint8_t *p_res5 = nullptr;                                                      // ... and the residues in the 5-bit transfer format
};

using namespace p7g;

extern "C" {

int p7gpu_dbfile_from_fasta(const char *fasta_path, const char *out_path, int nthreads, int align, int64_t *ret_nseq, int64_t *ret_nres)
{
  if (!fasta_path || !out_path || align < 1 || align > 4096 || (align & (align - 1))) return P7GPU_EINVAL;
  int fd = open(fasta_path, O_RDONLY);
  if (fd < 0) { set_error("dbfile_from_fasta: cannot open %s: %s", fasta_path, strerror(errno)); return P7GPU_EINVAL; }
  struct stat sb;
  if (fstat(fd, &sb) != 0) { close(fd); set_error("dbfile_from_fasta: fstat failed on %s", fasta_path); return P7GPU_ESYS; }
  const int64_t n = (int64_t) sb.st_size;
  const char *t = (const char *) "";
  void *map = nullptr;
  if (n > 0) {
    map = mmap(nullptr, (size_t) n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { close(fd); set_error("dbfile_from_fasta: mmap of %s failed: %s", fasta_path, strerror(errno)); return P7GPU_ESYS; }
    madvise(map, (size_t) n, MADV_SEQUENTIAL);
    t = (const char *) map;
  }
  auto done = [&](int code) { if (map) munmap(map, (size_t) n); close(fd); return code; };
  // the text before the first record may only be white space (header_fasta :3010-3017)
  {
    int64_t p = 0;
    while (p < n && isspace((unsigned char) t[p])) p++;
    if (p < n && t[p] != '>') { set_error("Line %lld: unexpected char %c; expected FASTA to start with >", (long long) line_of(t, p), t[p]); return done(P7GPU_EINVAL); }
  }
  unsigned hw = std::thread::hardware_concurrency();
  int T = nthreads > 0 ? nthreads : (int) std::min<unsigned>(hw ? hw : 1, 32);
  if ((int64_t) T > n / (1 << 16) + 1) T = (int) (n / (1 << 16) + 1);
  std::vector<Chunk> chunks((size_t) T);
  {
    std::vector<std::thread> th;
    for (int k = 0; k < T; k++) th.emplace_back([&, k]() { parse_range(t, n, n * k / T, n * (k + 1) / T, align, &chunks[(size_t) k]); });
    for (auto &x : th) x.join();
  }
  for (auto &c : chunks)
    if (c.err_pos >= 0) {
      const long long line = (long long) line_of(t, c.err_pos);
      if (c.err_kind == 1) set_error("Line %lld: illegal character %c", line, c.err_char);                 // seebuf, esl_sqio_ascii.c:2294
      else if (c.err_kind == 3) set_error("Line %lld: non-ASCII character in sequence", line);             // :2268
      else if (c.err_kind == 4) set_error("Premature EOF in parsing FASTA name/description line");         // :3047
      else if (c.err_kind == 5) set_error("Line %lld: target longer than 2^31-1 residues (unsupported)", line);
      else set_error("Line %lld: no FASTA name found", line);                                              // :3029
      return done(P7GPU_EINVAL);
    }
  // ---- assemble: residues (one leading sentinel), offsets, lengths, names ----
  const uint64_t lead = (uint64_t) align;                              // leading sentinels: the first target starts at `align` (1: one sentinel)
  uint64_t nseq = 0, resbytes = lead, namebytes = 0; int64_t nres = 0; int32_t Lmax = 0;
  for (auto &c : chunks) { nseq += c.len.size(); resbytes += c.res.size(); namebytes += c.names.size(); }
  std::vector<int64_t> offsets((size_t) nseq); std::vector<int32_t> lengths((size_t) nseq); std::vector<uint64_t> nameidx((size_t) nseq);
  {
    uint64_t s = 0, pos = lead, npos = 0;
    for (auto &c : chunks) {
      uint64_t q = pos;
      for (size_t j = 0; j < c.len.size(); j++, s++) {
        offsets[s] = (int64_t) q; lengths[s] = c.len[j]; q += ((uint64_t) c.len[j] + 1 + (uint64_t) align - 1) / (uint64_t) align * (uint64_t) align;
        nameidx[s] = npos + c.nameoff[j]; nres += c.len[j]; Lmax = std::max(Lmax, c.len[j]);
      }
      pos += c.res.size(); npos += c.names.size();
    }
  }
  int ofd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (ofd < 0) { set_error("dbfile_from_fasta: cannot create %s: %s", out_path, strerror(errno)); return done(P7GPU_EINVAL); }
  Header h; memset(&h, 0, sizeof(h));
  memcpy(h.magic, MAGIC, 8); h.version = 1; h.Kp = P7GPU_KP; h.nseq = nseq; h.nbytes = resbytes; h.nres = (uint64_t) nres; h.max_length = Lmax;
  uint64_t pos = 4096;
  h.off_residues = pos; pos += resbytes; pos = (pos + 4095) / 4096 * 4096;
  h.off_offsets = pos;  pos += 8 * nseq; pos = (pos + 4095) / 4096 * 4096;
  h.off_lengths = pos;  pos += 4 * nseq; pos = (pos + 4095) / 4096 * 4096;
  h.off_nameidx = pos;  pos += 8 * nseq; pos = (pos + 4095) / 4096 * 4096;
  h.off_names = pos; h.names_bytes = namebytes;
  bool bad = false;
  uint64_t w = 0;
  bad |= write_all(ofd, &h, sizeof(h)) != 0; w += sizeof(h); bad |= pad_to(ofd, &w, 4096) != 0;
  { std::vector<uint8_t> s0((size_t) lead, C_SENT); bad |= write_all(ofd, s0.data(), s0.size()) != 0; w += lead; }
  for (auto &c : chunks) { bad |= write_all(ofd, c.res.data(), c.res.size()) != 0; w += c.res.size(); }
  bad |= pad_to(ofd, &w, 4096) != 0;
  bad |= write_all(ofd, offsets.data(), 8 * (size_t) nseq) != 0; w += 8 * nseq; bad |= pad_to(ofd, &w, 4096) != 0;
  bad |= write_all(ofd, lengths.data(), 4 * (size_t) nseq) != 0; w += 4 * nseq; bad |= pad_to(ofd, &w, 4096) != 0;
  bad |= write_all(ofd, nameidx.data(), 8 * (size_t) nseq) != 0; w += 8 * nseq; bad |= pad_to(ofd, &w, 4096) != 0;
  for (auto &c : chunks) { bad |= write_all(ofd, c.names.data(), c.names.size()) != 0; w += c.names.size(); }
  bad |= pad_to(ofd, &w, 4096) != 0;
  bad |= close(ofd) != 0;
  if (bad) { set_error("dbfile_from_fasta: write to %s failed: %s", out_path, strerror(errno)); return done(P7GPU_ESYS); }
  if (ret_nseq) *ret_nseq = (int64_t) nseq;
  if (ret_nres) *ret_nres = nres;
  return done(P7GPU_OK);
}

int p7gpu_dbfile_open(const char *path, p7gpu_dbfile **ret)
{
  if (!path || !ret) return P7GPU_EINVAL;
  *ret = nullptr;
  int fd = open(path, O_RDONLY);
  if (fd < 0) { set_error("dbfile_open: cannot open %s: %s", path, strerror(errno)); return P7GPU_EINVAL; }
  struct stat sb;
  if (fstat(fd, &sb) != 0 || (size_t) sb.st_size < sizeof(Header)) { close(fd); set_error("dbfile_open: %s is not a packed database", path); return P7GPU_EINVAL; }
  void *map = mmap(nullptr, (size_t) sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
  if (map == MAP_FAILED) { close(fd); set_error("dbfile_open: mmap of %s failed: %s", path, strerror(errno)); return P7GPU_ESYS; }
  const Header *h = (const Header *) map;
  const uint64_t sz = (uint64_t) sb.st_size;
  // every (offset, size) pair is checked against the file size without forming a sum or product that could wrap
  auto fits = [sz](uint64_t off, uint64_t bytes) { return off <= sz && bytes <= sz - off && (off & 4095) == 0; };
  bool ok = memcmp(h->magic, MAGIC, 8) == 0 && h->version == 1 && h->Kp == P7GPU_KP && h->max_length >= 0 &&
            h->nseq <= sz / 8 && h->nseq <= 0x7fffffffffULL && h->nbytes >= 1 &&
            fits(h->off_residues, h->nbytes) && fits(h->off_offsets, 8 * h->nseq) && fits(h->off_lengths, 4 * h->nseq) &&
            fits(h->off_nameidx, 8 * h->nseq) && fits(h->off_names, h->names_bytes);
  if (ok && h->nseq > 0) {                                            // the names section ends in a NUL: no accessor can run off the mapping
    const char *names = (const char *) map + h->off_names;
    ok = h->names_bytes >= 2 && names[h->names_bytes - 1] == 0;
  }
  // (per-target offsets and lengths are validated by the search itself, chunk by chunk while earlier chunks compute: stream_msv;
  //  per-target name indices by p7gpu_dbfile_name/desc — neither costs a pass over a 10^7-entry index at open time)
  if (!ok) { munmap(map, (size_t) sb.st_size); close(fd); set_error("dbfile_open: %s: bad magic, version, section table or index (truncated or corrupt file)", path); return P7GPU_EINVAL; }
  p7gpu_dbfile *d = new p7gpu_dbfile();
  d->map = map; d->map_bytes = (size_t) sb.st_size; d->fd = fd; d->h = h;
  *ret = d;
  return P7GPU_OK;
}

void p7gpu_dbfile_close(p7gpu_dbfile *d)
{
  if (!d) return;
  if (d->p_res) cudaFreeHost(d->p_res);
  if (d->p_res5) cudaFreeHost(d->p_res5);
  if (d->p_off) cudaFreeHost(d->p_off);
  if (d->p_len) cudaFreeHost(d->p_len);
  if (d->map) munmap(d->map, d->map_bytes);
  if (d->fd >= 0) close(d->fd);
  delete d;
}

int64_t        p7gpu_dbfile_nseq     (const p7gpu_dbfile *d) { return d ? (int64_t) d->h->nseq : 0; }
int64_t        p7gpu_dbfile_nres     (const p7gpu_dbfile *d) { return d ? (int64_t) d->h->nres : 0; }
int64_t        p7gpu_dbfile_nbytes   (const p7gpu_dbfile *d) { return d ? (int64_t) d->h->nbytes : 0; }
int32_t        p7gpu_dbfile_maxlength(const p7gpu_dbfile *d) { return d ? d->h->max_length : 0; }
const uint8_t *p7gpu_dbfile_residues (const p7gpu_dbfile *d) { return !d ? nullptr : d->pinned ? d->p_res : (const uint8_t *) d->map + d->h->off_residues; }
const int64_t *p7gpu_dbfile_offsets  (const p7gpu_dbfile *d) { return !d ? nullptr : d->pinned ? d->p_off : (const int64_t *) ((const char *) d->map + d->h->off_offsets); }
const int32_t *p7gpu_dbfile_lengths  (const p7gpu_dbfile *d) { return !d ? nullptr : d->pinned ? d->p_len : (const int32_t *) ((const char *) d->map + d->h->off_lengths); }
const char *p7gpu_dbfile_name(const p7gpu_dbfile *d, int64_t s)
{
  if (!d || s < 0 || (uint64_t) s >= d->h->nseq) return nullptr;
  const uint64_t *idx = (const uint64_t *) ((const char *) d->map + d->h->off_nameidx);
  if (idx[s] >= d->h->names_bytes) { set_error("dbfile_name: corrupt name index for target %lld", (long long) s); return nullptr; }
  return (const char *) d->map + d->h->off_names + idx[s];          // NUL-terminated inside the section (checked at open)
}
const char *p7gpu_dbfile_desc(const p7gpu_dbfile *d, int64_t s)
{
  const char *nm = p7gpu_dbfile_name(d, s);
  if (!nm) return nullptr;
  const char *end = (const char *) d->map + d->h->off_names + d->h->names_bytes;
  const char *ds = nm + strlen(nm) + 1;
  return ds < end ? ds : end - 1;                                    // a record cut off after its name: empty description
}

// Load the database into page-locked host memory so that the streamed upload runs at full PCIe rate (optional, once per open;
// a read-only file mapping cannot be registered with the driver, so this is a copy, made by all host threads).
int p7gpu_dbfile_pin(p7gpu_dbfile *d)
{
  if (!d) return P7GPU_EINVAL;
  if (d->pinned) return P7GPU_OK;
  const size_t nb = (size_t) d->h->nbytes, ns = (size_t) d->h->nseq;
  P7G_CUDA(cudaHostAlloc((void **) &d->p_res, std::max<size_t>(nb, 1), cudaHostAllocPortable));
  P7G_CUDA(cudaHostAlloc((void **) &d->p_off, std::max<size_t>(8 * ns, 8), cudaHostAllocPortable));
  P7G_CUDA(cudaHostAlloc((void **) &d->p_len, std::max<size_t>(4 * ns, 4), cudaHostAllocPortable));
  const uint8_t *src = (const uint8_t *) d->map + d->h->off_residues;
  unsigned hw = std::thread::hardware_concurrency();
  const size_t T = std::max<size_t>(1, std::min<size_t>(hw ? hw : 1, 16));
  std::vector<std::thread> th;
  for (size_t k = 0; k < T; k++) th.emplace_back([=]() { const size_t lo = nb * k / T, hi = nb * (k + 1) / T; memcpy(d->p_res + lo, src + lo, hi - lo); });
  for (auto &x : th) x.join();
  memcpy(d->p_off, (const char *) d->map + d->h->off_offsets, 8 * ns);
  memcpy(d->p_len, (const char *) d->map + d->h->off_lengths, 4 * ns);
  // the streamed search sends 5/8 of the bytes when it has the 5-bit form (p7gpu_pipeline_run_host5)
  P7G_CUDA(cudaHostAlloc((void **) &d->p_res5, (size_t) p7gpu_pack5_bytes((int64_t) nb), cudaHostAllocPortable));
  P7G_CHECK(p7gpu_pack5(d->p_res, (int64_t) nb, d->p_res5, 0));
  d->pinned = true;
  return P7GPU_OK;
}

int p7gpu_pipeline_run_file(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_dbfile *d, const p7gpu_pipeline_opts *opts,
                            p7gpu_pipeline_result **ret_res)
{
  if (!d) return P7GPU_EINVAL;
  if (d->pinned && d->p_res5)
    return p7gpu_pipeline_run_host5(ctx, prof, d->p_res5, (int64_t) d->h->nbytes, d->p_off, d->p_len, (int64_t) d->h->nseq, opts, ret_res);
  return p7gpu_pipeline_run_host(ctx, prof, p7gpu_dbfile_residues(d), (int64_t) d->h->nbytes, p7gpu_dbfile_offsets(d),
                                 p7gpu_dbfile_lengths(d), (int64_t) d->h->nseq, opts, ret_res);
}

int p7gpu_pipeline_run_file_sharded(p7gpu_ctx *const *ctxs, const p7gpu_profile *const *profs, int nshards, const p7gpu_dbfile *d,
                                    const p7gpu_pipeline_opts *opts, int64_t *first, p7gpu_pipeline_result **ret_res)
{
  if (!d) return P7GPU_EINVAL;
  const bool five = d->pinned && d->p_res5;
  return p7gpu_pipeline_run_sharded(ctxs, profs, nshards, five ? nullptr : p7gpu_dbfile_residues(d), five ? d->p_res5 : nullptr,
                                    (int64_t) d->h->nbytes, p7gpu_dbfile_offsets(d), p7gpu_dbfile_lengths(d), (int64_t) d->h->nseq,
                                    opts, first, ret_res);
}

}  // extern "C"
