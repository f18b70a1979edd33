// lra_b200 -- the host program above the C ABI (include/lra_b200.h), with the reference's command line for the path this library replaces:
//
//   lra_b200 index  -CCS|-CLR|-ONT|-CONTIG ref.fa                      (lra.cpp:1041-1045 -> RunStoreIndex :780-995: writes ref.fa.mms / ref.fa.gli)
//   lra_b200 align  -CLR|-ONT ref.fa reads.fa|reads.fq [-t N] [-p s|p|pc|b] [-o out] [-a] [--PrintNumAln n] [--batch-bases n] [--device d]
//                                                                      (lra.cpp:1056-1059 -> RunAlign :174-728)
//
// Plain C++ (no CUDA, no torch): everything on the device goes through liblra_b200.so.  What it keeps of the reference's host side:
//   * Genome::Read (Genome.h:115-138): contigs upper-cased, name = first token of the header line;
//   * Input::GetNext for FASTA / FASTQ (Input.h:182-290), plain or gzip: name = first token, blanks dropped, bases upper-cased, several files chained;
//     FASTA is parsed on all host threads;
//   * ReadIndex / LocalIndex::Read (MMIndex.h:154-173, 402-412): the on-disk .mms / .gli formats, globalK taken from the .mms;
//   * the SAM header (@PG line, Header::WriteSAMHeader, Genome.h:85-89) and the printers through lra_b200_format_records.
// Reads are mapped in batches of --batch-bases bases; records come back in input order (the reference's order with -t 1).
// -t only sets the number of host threads that format text (the GPU replaces the reference's pthreads workers).
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>
#include "../../include/lra_b200.h"

namespace {

struct Fasta {                      // sequences back to back + names + offsets
  std::string seq;
  std::string qual;                 // FASTQ: quality strings at the same offsets (blanks dropped); empty for FASTA
  std::vector<std::string> names;
  std::vector<uint64_t> off;        // n + 1 entries
};

// whole file into memory; gzip-compressed files are inflated through zlib (gzread also passes plain files through, as the reference's kseq does)
bool slurp(const std::string &path, std::string &out) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) return false;
  gzbuffer(f, 1 << 20);
  out.clear();
  std::vector<char> buf(8u << 20);
  for (;;) {
    const int n = gzread(f, buf.data(), (unsigned)buf.size());
    if (n < 0) { gzclose(f); return false; }
    if (n == 0) break;
    out.append(buf.data(), (size_t)n);
  }
  gzclose(f);
  return true;
}

std::string first_token(const char *p, const char *e) {   // after the leading '>' / '@'
  while (p < e && isspace((unsigned char)*p)) p++;
  const char *q = p;
  while (q < e && !isspace((unsigned char)*q)) q++;
  return std::string(p, q);
}

// FASTA or FASTQ text -> records appended to `fa`; keep_case_blank = false applies the read rules (drop blanks, upper-case)
bool parse_reads(const std::string &text, Fasta &fa) {
  const char *p = text.data(), *e = p + text.size();
  if (fa.off.empty()) fa.off.push_back(0);
  while (p < e) {
    const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
    const char *le = nl ? nl : e;
    if (*p == '>') {
      fa.names.push_back(first_token(p + 1, le));
      p = nl ? nl + 1 : e;
      while (p < e && *p != '>') {
        nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        le = nl ? nl : e;
        for (const char *c = p; c < le; c++) if (*c != ' ' && *c != '\r') fa.seq.push_back((char)toupper((unsigned char)*c));
        p = nl ? nl + 1 : e;
      }
      fa.off.push_back(fa.seq.size());
    } else if (*p == '@') {
      fa.names.push_back(first_token(p + 1, le));
      p = nl ? nl + 1 : e;
      nl = (const char *)memchr(p, '\n', (size_t)(e - p)); le = nl ? nl : e;
      for (const char *c = p; c < le; c++) if (*c != ' ' && *c != '\r') fa.seq.push_back((char)toupper((unsigned char)*c));
      p = nl ? nl + 1 : e;
      nl = (const char *)memchr(p, '\n', (size_t)(e - p)); p = nl ? nl + 1 : e;                                           // '+' line
      nl = (const char *)memchr(p, '\n', (size_t)(e - p)); le = nl ? nl : e;
      fa.qual.resize(fa.off.back(), '!');                                                                                  // (earlier FASTA records of a mixed input)
      for (const char *c = p; c < le; c++) if (*c != ' ' && *c != '\r') fa.qual.push_back(*c);
      fa.qual.resize(fa.seq.size(), '!');
      p = nl ? nl + 1 : e;
      fa.off.push_back(fa.seq.size());
    } else {
      p = nl ? nl + 1 : e;      // blank line
    }
  }
  return true;
}

// the same on host threads: the text is cut at record starts ('>' at a line start; FASTQ is parsed by one thread because '@' is also a
// quality character), the pieces are parsed independently and concatenated in order
bool parse_reads_mt(const std::string &text, Fasta &fa, int threads) {
  if (text.empty()) { if (fa.off.empty()) fa.off.push_back(0); return true; }
  if (threads < 2 || text.size() < (8u << 20) || text[0] != '>') return parse_reads(text, fa);
  std::vector<size_t> cut(1, 0);
  for (int t = 1; t < threads; t++) {
    size_t p = text.size() * (size_t)t / (size_t)threads;
    const char *q = nullptr;
    while (p < text.size() && (q = (const char *)memchr(text.data() + p, '>', text.size() - p))) {
      p = (size_t)(q - text.data());
      if (p == 0 || text[p - 1] == '\n') break;
      p++;
    }
    if (!q || p >= text.size()) break;
    if (p > cut.back()) cut.push_back(p);
  }
  cut.push_back(text.size());
  const int np = (int)cut.size() - 1;
  std::vector<Fasta> part(np);
  std::vector<std::thread> th;
  for (int i = 0; i < np; i++) th.emplace_back([&, i] { std::string piece(text.data() + cut[i], cut[i + 1] - cut[i]); parse_reads(piece, part[i]); });
  for (auto &x : th) x.join();
  if (fa.off.empty()) fa.off.push_back(0);
  for (int i = 0; i < np; i++) {
    const uint64_t base = fa.seq.size();
    fa.seq += part[i].seq;
    if (!part[i].qual.empty()) { fa.qual.resize(base, '!'); fa.qual += part[i].qual; }
    for (size_t r = 0; r < part[i].names.size(); r++) { fa.names.push_back(std::move(part[i].names[r])); fa.off.push_back(base + part[i].off[r + 1]); }
  }
  return true;
}

// Genome::Read: kseq semantics (sequence = all non-blank characters of the record's lines), upper-cased
bool parse_genome(const std::string &text, Fasta &fa) { return parse_reads_mt(text, fa, (int)std::thread::hardware_concurrency()); }

struct Preset { int k, w, max_freq, win, per_window; };
bool index_preset(const std::string &m, Preset &p) {      // lra.cpp:884-911
  if (m == "-ONT" || m == "-CCS") { p = {17, 10, 150, 15, 1}; return true; }
  if (m == "-CLR") { p = {15, 10, 250, 12, 1}; return true; }
  if (m == "-CONTIG") { p = {19, 10, 30, 20, 1}; return true; }
  return false;
}

int die(lra_b200_ctx *ctx, const char *what) { fprintf(stderr, "lra_b200: %s: %s\n", what, lra_b200_last_error(ctx)); return 1; }

void write_header(FILE *f, const Fasta &g) {              // Header::Write (Genome.h:45-60)
  const int32_t nc = (int32_t)g.names.size();
  fwrite(&nc, 4, 1, f);
  for (const std::string &n : g.names) { const int32_t l = (int32_t)n.size(); fwrite(&l, 4, 1, f); fwrite(n.data(), 1, n.size(), f); }
  fwrite(g.off.data(), 8, g.off.size(), f);
}

int run_index(int argc, char **argv) {
  std::string mode = "-ONT", ref; int device = 0;
  for (int i = 0; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "-ONT" || a == "-CCS" || a == "-CLR" || a == "-CONTIG") mode = a;
    else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
    else if (a[0] != '-') ref = a;
  }
  Preset pr;
  if (ref.empty() || !index_preset(mode, pr)) { fprintf(stderr, "usage: lra_b200 index -CCS|-CLR|-ONT|-CONTIG ref.fa\n"); return 1; }
  std::string text; Fasta g;
  if (!slurp(ref, text) || !parse_genome(text, g) || g.names.empty()) { fprintf(stderr, "Cannot open target %s\n", ref.c_str()); return 1; }
  text.clear(); text.shrink_to_fit();
  lra_b200_ctx *ctx = nullptr;
  if (lra_b200_create(&ctx, device)) return die(nullptr, "no CUDA device");
  const int nc = (int)g.names.size();
  std::vector<uint64_t> start(g.off.begin(), g.off.end() - 1); std::vector<uint32_t> len(nc);
  for (int c = 0; c < nc; c++) len[c] = (uint32_t)(g.off[c + 1] - g.off[c]);
  lra_b200_seq *arena = nullptr;
  if (lra_b200_seq_upload(ctx, g.seq.data(), g.seq.size(), &arena)) return die(ctx, "genome upload");
  // global index
  lra_b200_index *ix = nullptr;
  if (lra_b200_gindex_build(ctx, arena, start.data(), len.data(), nc, pr.k, pr.w, pr.max_freq, pr.win, pr.per_window, &ix)) return die(ctx, "global index");
  const uint64_t n = lra_b200_index_size(ix);
  std::vector<uint64_t> t(n); std::vector<uint32_t> pos(n);
  if (lra_b200_index_download(ctx, ix, t.data(), pos.data())) return die(ctx, "index download");
  lra_b200_index_free(ctx, ix);
  {
    FILE *f = fopen((ref + ".mms").c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s.mms\n", ref.c_str()); return 1; }
    const int64_t n64 = (int64_t)n; const int32_t k32 = pr.k;
    fwrite(&n64, 8, 1, f); fwrite(&k32, 4, 1, f);
    write_header(f, g);
    struct Rec { uint64_t t; uint32_t pos, pad; };
    std::vector<Rec> buf(1 << 20);
    for (uint64_t a = 0; a < n; a += buf.size()) {
      const uint64_t m = std::min<uint64_t>(buf.size(), n - a);
      for (uint64_t i = 0; i < m; i++) { buf[i].t = t[a + i]; buf[i].pos = pos[a + i]; buf[i].pad = 0; }
      fwrite(buf.data(), sizeof(Rec), m, f);
    }
    fclose(f);
  }
  fprintf(stderr, "There are %llu minimizers left\n", (unsigned long long)n);
  // local index (k 10, w 5, window 2048, maxFreq 15: MMIndex.h:110-126, lra.cpp:795-818)
  lra_b200_lindex *li = nullptr;
  if (lra_b200_lindex_build(ctx, arena, start.data(), len.data(), nc, 10, 5, 2048, 15, &li)) return die(ctx, "local index");
  uint64_t nw = 0, nm = 0;
  lra_b200_lindex_sizes(li, &nw, &nm);
  std::vector<uint64_t> wo(nw + 1), bd(nw + 1); std::vector<uint32_t> mn(nm ? nm : 1);
  if (lra_b200_lindex_download(ctx, li, wo.data(), bd.data(), mn.data())) return die(ctx, "local index download");
  {
    FILE *f = fopen((ref + ".gli").c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s.gli\n", ref.c_str()); return 1; }
    const int32_t h[4] = {10, 5, 2048, (int32_t)(nw + 1)};
    fwrite(h, 4, 4, f); fwrite(wo.data(), 8, nw + 1, f); fwrite(bd.data(), 8, nw + 1, f);
    const uint64_t nm64 = nm; fwrite(&nm64, 8, 1, f); fwrite(mn.data(), 4, nm, f);
    fclose(f);
  }
  lra_b200_lindex_free(ctx, li);
  lra_b200_seq_free(ctx, arena);
  lra_b200_destroy(ctx);
  return 0;
}

bool read_mms(const std::string &path, int &k, std::vector<uint64_t> &t, std::vector<uint32_t> &pos) {
  std::string d;
  if (!slurp(path, d) || d.size() < 16) return false;
  int64_t n; int32_t nc;
  memcpy(&n, d.data(), 8); memcpy(&k, d.data() + 8, 4); memcpy(&nc, d.data() + 12, 4);
  size_t o = 16;
  for (int c = 0; c < nc; c++) { int32_t l; memcpy(&l, d.data() + o, 4); o += 4 + (size_t)l; }
  o += 8 * (size_t)(nc + 1);
  if (o + 16ull * (uint64_t)n > d.size()) return false;
  t.resize((size_t)n); pos.resize((size_t)n);
  for (int64_t i = 0; i < n; i++) { memcpy(&t[i], d.data() + o + 16 * i, 8); memcpy(&pos[i], d.data() + o + 16 * i + 8, 4); }
  return true;
}

bool read_gli(const std::string &path, int &k, int &w, int &window, std::vector<uint64_t> &so, std::vector<uint64_t> &tb, std::vector<uint32_t> &mn) {
  std::string d;
  if (!slurp(path, d) || d.size() < 24) return false;
  int32_t h[4]; memcpy(h, d.data(), 16);
  k = h[0]; w = h[1]; window = h[2];
  const size_t n = (size_t)h[3];
  so.resize(n); tb.resize(n);
  memcpy(so.data(), d.data() + 16, 8 * n); memcpy(tb.data(), d.data() + 16 + 8 * n, 8 * n);
  uint64_t nm; memcpy(&nm, d.data() + 16 + 16 * n, 8);
  mn.resize((size_t)nm);
  memcpy(mn.data(), d.data() + 24 + 16 * n, 4 * (size_t)nm);
  return true;
}

int run_align(int argc, char **argv, const std::string &cmdline) {
  std::string mode = "-ONT", ref, out_path, fmt = "p";      // the reference's default printer is PAF (Options.h:162)
  std::vector<std::string> inputs;
  int device = 0, print_num = -1, print_md = 0; uint64_t batch_bases = 256ull << 20;
  for (int i = 0; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "-ONT" || a == "-CLR" || a == "-CCS" || a == "-CONTIG") mode = a;
    else if (a == "-t" && i + 1 < argc) setenv("LRA_B200_SAM_THREADS", argv[++i], 1);
    else if (a == "-p" && i + 1 < argc) fmt = argv[++i];
    else if (a == "-o" && i + 1 < argc) out_path = argv[++i];
    else if (a == "--PrintNumAln" && i + 1 < argc) print_num = atoi(argv[++i]);
    else if (a == "--printMD") print_md = 1;
    else if (a == "--batch-bases" && i + 1 < argc) batch_bases = strtoull(argv[++i], nullptr, 10);
    else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
    else if (a[0] != '-') { if (ref.empty()) ref = a; else inputs.push_back(a); }
  }
  if (ref.empty() || inputs.empty()) { fprintf(stderr, "usage: lra_b200 align -CLR|-ONT ref.fa reads.fa [-t N] [-p s|p|pc|b|a] [--printMD] [-o out]\n"); return 1; }
  const char f = fmt == "s" ? 's' : fmt == "p" ? 'p' : fmt == "pc" ? 'c' : fmt == "b" ? 'b' : fmt == "a" ? 'a' : 0;
  if (!f) { fprintf(stderr, "lra_b200 align: -p %s is not supported (s, p, pc, b, a)\n", fmt.c_str()); return 1; }
  std::string text; Fasta g;
  if (!slurp(ref, text) || !parse_genome(text, g) || g.names.empty()) { fprintf(stderr, "Cannot open target %s\n", ref.c_str()); return 1; }
  text.clear(); text.shrink_to_fit();
  int gk = 0, lk = 0, lw = 0, lwin = 0;
  std::vector<uint64_t> mt, so, tb; std::vector<uint32_t> mp, mn;
  if (!read_mms(ref + ".mms", gk, mt, mp) || !read_gli(ref + ".gli", lk, lw, lwin, so, tb, mn)) {
    fprintf(stderr, "lra_b200 align: cannot read %s.mms / .gli (run `lra_b200 index %s %s` or the reference's `lra index`)\n", ref.c_str(), mode.c_str(), ref.c_str());
    return 1;
  }
  lra_b200_ctx *ctx = nullptr;
  if (lra_b200_create(&ctx, device)) return die(nullptr, "no CUDA device");
  lra_b200_map_opts opts;
  if (lra_b200_map_opts_preset(mode.c_str(), &opts)) return die(ctx, "preset");
  opts.globalK = gk; opts.smallK = lk; opts.smallW = lw; opts.localIndexWindow = lwin;
  if (print_num > 0) opts.PrintNumAln = print_num;
  const int nc = (int)g.names.size();
  lra_b200_mapper *m = nullptr;
  if (lra_b200_mapper_create(ctx, &opts, g.seq.data(), g.seq.size(), g.off.data(), nc, mt.data(), mp.data(), mt.size(), so.data(), tb.data(), (int32_t)so.size(), mn.data(),
                             mn.size(), &m))
    return die(ctx, "mapper_create");
  { std::vector<uint64_t>().swap(mt); std::vector<uint32_t>().swap(mp); std::vector<uint32_t>().swap(mn); }
  FILE *out = out_path.empty() || out_path == "-" ? stdout : fopen(out_path.c_str(), "wb");
  if (!out) { fprintf(stderr, "cannot write %s\n", out_path.c_str()); return 1; }
  std::string cnames; std::vector<uint64_t> clen(nc);
  for (int c = 0; c < nc; c++) { cnames += g.names[c]; cnames.push_back('\0'); clen[c] = g.off[c + 1] - g.off[c]; }
  if (f == 's') {
    fprintf(out, "@PG\tID:lra\tPN:lra\tVN:1.3.7.1\tCL:%s\n", cmdline.c_str());
    for (int c = 0; c < nc; c++) fprintf(out, "@SQ\tSN:%s\tLN:%llu\n", g.names[c].c_str(), (unsigned long long)clen[c]);
  }
  Fasta rd;
  for (const std::string &in : inputs) {
    std::string t2;
    if (!slurp(in, t2)) { fprintf(stderr, "Cannot open reads %s\n", in.c_str()); return 1; }
    parse_reads_mt(t2, rd, (int)std::thread::hardware_concurrency());
  }
  const size_t n_reads = rd.names.size();
  std::vector<char> textbuf;
  unsigned long long mapped = 0;
  for (size_t r0 = 0; r0 < n_reads;) {
    size_t r1 = r0; uint64_t bases = 0;
    while (r1 < n_reads && (r1 == r0 || bases + (rd.off[r1 + 1] - rd.off[r1]) <= batch_bases)) { bases += rd.off[r1 + 1] - rd.off[r1]; r1++; }
    const int n = (int)(r1 - r0);
    std::vector<uint64_t> off(n); std::vector<uint32_t> len(n);
    for (int i = 0; i < n; i++) { off[i] = rd.off[r0 + i] - rd.off[r0]; len[i] = (uint32_t)(rd.off[r0 + i + 1] - rd.off[r0 + i]); }
    lra_b200_map_result res; memset(&res, 0, sizeof res);
    std::vector<int32_t> status(n), n_aln(n), nseg(4 * n), seg0(4 * n), rank(4 * n);
    std::vector<lra_b200_record> recs((size_t)3 * n + 1024 + 65536); std::vector<uint32_t> cig((size_t)bases + 64ull * n + 4096);
    res.status = status.data(); res.n_aln = n_aln.data(); res.aln_nseg = nseg.data(); res.aln_seg0 = seg0.data(); res.aln_rank = rank.data();
    res.records = recs.data(); res.record_cap = recs.size(); res.cigar = cig.data(); res.cigar_cap = cig.size();
    const char *base = rd.seq.data() + rd.off[r0];
    if (lra_b200_map_batch(ctx, m, base, bases, off.data(), len.data(), n, &res)) return die(ctx, "map_batch");
    std::string names;
    for (int i = 0; i < n; i++) { names += rd.names[r0 + i]; names.push_back('\0'); if (status[i] == 0 && n_aln[i] > 0) mapped++; }
    const char *qbase = rd.qual.size() == rd.seq.size() && !rd.qual.empty() ? rd.qual.data() + rd.off[r0] : nullptr;
    int64_t need = lra_b200_format_records_ref(&opts, &res, n, names.data(), base, qbase, off.data(), len.data(), cnames.data(), clen.data(), nc, g.seq.data(), g.off.data(), f, print_md, 0, nullptr, 0);
    need = -need;
    textbuf.resize((size_t)need + 16);
    const int64_t got = lra_b200_format_records_ref(&opts, &res, n, names.data(), base, qbase, off.data(), len.data(), cnames.data(), clen.data(), nc, g.seq.data(), g.off.data(), f, print_md, 0,
                                                    textbuf.data(), need + 16);
    if (got > 0) fwrite(textbuf.data(), 1, (size_t)got, out);
    r0 = r1;
  }
  if (out != stdout) fclose(out);
  fprintf(stderr, "lra_b200 align: %zu reads, %llu with an alignment\n", n_reads, mapped);
  lra_b200_mapper_destroy(ctx, m);
  lra_b200_destroy(ctx);
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: lra_b200 index|align ... (the command line of `lra`, see the header of lra_b200/cli/lra_b200_cli.cpp)\n"); return 1; }
  std::string cmdline = "lra align";
  for (int i = 0; i < argc; i++) { cmdline += " "; cmdline += argv[i]; }
  const std::string cmd = argv[1];
  if (cmd == "index" || cmd == "global") return run_index(argc - 2, argv + 2);
  if (cmd == "align") return run_align(argc - 2, argv + 2, cmdline);
  fprintf(stderr, "lra_b200: unknown command %s\n", cmd.c_str());
  return 1;
}
