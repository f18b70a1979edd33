// Host side of the MapRead seam (included at the end of lra_b200.cu): the align presets, lra_b200_mapper_create / lra_b200_map_batch /
// lra_b200_mapper_destroy (packs the reads, builds their LocalIndex images, runs the mapper worker kernel, the a19 / a21 batch kernels and the
// finalize kernel, copies the records home) and the SAM emitter (Alignment::PrintSAM, host code).
#pragma once
#include <sstream>
#include <thread>
#include <algorithm>
#include "mp_map.cuh"

// ---- presets (lra.cpp:339-431 on top of Options.h:123-240) -------------------------------------------------------------------------------------
extern "C" int lra_b200_map_opts_preset(const char *mode, lra_b200_map_opts *o) {
  if (!mode || !o) return LRA_B200_EINVAL;
  memset(o, 0, sizeof *o);
  // defaults
  o->globalK = 17; o->globalW = 10; o->globalMaxFreq = 50; o->localW = 5; o->localMaxFreq = 30; o->smallK = 10; o->smallW = 5;
  o->cleanMaxDiag = 100; o->minDiagCluster = 10; o->cleanClustersize = 100; o->SecondCleanMinDiagCluster = 40; o->SecondCleanMaxDiag = 10; o->punish_anchorfreq = 10;
  o->anchorPerlength = 10; o->NumAln = 3; o->PrintNumAln = 1; o->splitdist = 50000; o->readType = 0;
  o->initial_anchorbonus = 1.0f; o->second_anchorbonus = 2.0f; o->alnthres = 0.7f; o->anchorstoosparse = 0.01f;
  o->refineSpaceDist = 10000; o->window = 100; o->limitrefine = 1; o->RefineBySDP = 1;
  o->localMatch = 4; o->localMismatch = -3; o->localIndel = -4; o->localBand = 15; o->refineBand = 7; o->hardClip = 0; o->bypassClustering = 0;
  o->gapopen = 2.0f; o->gapextend = 10.0f; o->gaproot = 2.0f; o->gapCeiling1 = 1500; o->gapCeiling2 = 3000; o->localIndexWindow = 2048; o->localIndexMaxFreq = 15;
  o->HighlyAccurate = 0; o->maxDiag = 500; o->maxGap = 5000; o->RoughClustermaxGap = 1000; o->minClusterSize = 2; o->minUniqueStretchNum = 1; o->minUniqueStretchDist = 50;
  o->merge_dist = 100;
  std::string m(mode);
  if (!m.empty() && m[0] == '-') m = m.substr(1);
  for (auto &c : m) c = (char)toupper(c);
  if (m == "CLR") {
    o->globalK = 15; o->globalW = 10; o->globalMaxFreq = 250; o->localW = 5; o->localMaxFreq = 15; o->readType = 1; o->refineBand = 20;
    o->gaproot = 1.5f; o->gapextend = 10.0f; o->gapopen = 7.0f; o->initial_anchorbonus = 15.0f; o->localMismatch = -1; o->localIndel = -2;
    o->gapCeiling1 = 1500; o->gapCeiling2 = 3000; o->NumAln = 2; o->PrintNumAln = 1; o->cleanMaxDiag = 200; o->SecondCleanMaxDiag = 120; o->SecondCleanMinDiagCluster = 10;
    o->refineSpaceDist = 30000; o->minDiagCluster = 3; o->bypassClustering = 1; o->punish_anchorfreq = 5; o->anchorPerlength = 5; o->cleanClustersize = 100;
    o->anchorstoosparse = 0.005f; o->hardClip = 1; o->alnthres = 0.50f; o->second_anchorbonus = 6.0f;
  } else if (m == "ONT") {
    o->globalK = 17; o->globalW = 10; o->globalMaxFreq = 150; o->localW = 5; o->localMaxFreq = 15; o->readType = 0;
    o->gaproot = 1.5f; o->gapextend = 10.0f; o->gapopen = 7.0f; o->initial_anchorbonus = 20.0f; o->localMismatch = -1; o->localIndel = -2;
    o->gapCeiling1 = 1500; o->gapCeiling2 = 3000; o->NumAln = 2; o->PrintNumAln = 1; o->cleanMaxDiag = 200; o->SecondCleanMaxDiag = 100; o->SecondCleanMinDiagCluster = 10;
    o->refineSpaceDist = 30000; o->minDiagCluster = 3; o->bypassClustering = 1; o->punish_anchorfreq = 5; o->anchorPerlength = 5; o->cleanClustersize = 100;
    o->anchorstoosparse = 0.005f; o->hardClip = 1; o->alnthres = 0.65f;
  } else if (m == "CCS") {
    o->globalK = 25; o->globalW = 20; o->globalMaxFreq = 150; o->localMaxFreq = 15; o->readType = 2;
    o->gaproot = 1.5f; o->gapextend = 15.0f; o->gapopen = 4.0f; o->initial_anchorbonus = 10.0f; o->gapCeiling1 = 2000; o->gapCeiling2 = 3000; o->HighlyAccurate = 1;
    o->NumAln = 2; o->PrintNumAln = 1; o->merge_dist = 100; o->RoughClustermaxGap = 500; o->maxGap = 400; o->cleanMaxDiag = 150; o->SecondCleanMaxDiag = 100;
    o->SecondCleanMinDiagCluster = 30; o->minDiagCluster = 10; o->minClusterSize = 10; o->cleanClustersize = 100; o->punish_anchorfreq = 10; o->anchorPerlength = 10;
    o->refineSpaceDist = 30000; o->anchorstoosparse = 0.005f; o->hardClip = 1;
  } else if (m == "CONTIG") {
    o->globalK = 19; o->globalW = 10; o->globalMaxFreq = 30; o->localMaxFreq = 15; o->readType = 3; o->refineBand = 50;
    o->gaproot = 1.5f; o->gapextend = 20.0f; o->gapopen = 4.0f; o->gapCeiling1 = 3000; o->gapCeiling2 = 5000; o->HighlyAccurate = 1; o->initial_anchorbonus = 1.0f;
    o->maxDiag = 100; o->maxGap = 500; o->RoughClustermaxGap = 500; o->NumAln = 2; o->PrintNumAln = 1; o->anchorstoosparse = 0.005f; o->merge_dist = 100;
    o->cleanMaxDiag = 150; o->SecondCleanMaxDiag = 100; o->SecondCleanMinDiagCluster = 30; o->minDiagCluster = 30; o->minClusterSize = 10; o->refineSpaceDist = 50000;
    o->cleanClustersize = 100; o->punish_anchorfreq = 10; o->anchorPerlength = 10; o->hardClip = 1;
  } else return LRA_B200_EINVAL;
  return LRA_B200_OK;
}

// ---- SAM emitter (Alignment.h:658-808, 811-832; Mapping_ultility.h:457-494) -------------------------------------------------------------------
static const unsigned char *mp_revcomp_table() {
  static unsigned char T[256]; static bool init = false;
  if (!init) { for (int i = 0; i < 256; i++) T[i] = 'N'; T['A'] = 'T'; T['C'] = 'G'; T['G'] = 'C'; T['T'] = 'A'; T['a'] = 't'; T['c'] = 'g'; T['g'] = 'c'; T['t'] = 'a'; T['n'] = 'n'; init = true; }
  return T;
}

// appender with the ostream formatting the reference relies on (operator<< of integers, chars, strings and one float with the default precision)
struct SamOut {
  std::string s;
  void put(const char *p, size_t n) { s.append(p, n); }
  void put(const char *p) { s.append(p); }
  void put(char c) { s.push_back(c); }
  void u(unsigned long long v) { char b[24]; int n = 0; do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v); while (n) s.push_back(b[--n]); }
  void i(long long v) { if (v < 0) { s.push_back('-'); u((unsigned long long)(-(v + 1)) + 1ull); } else u((unsigned long long)v); }
  void f(float v) { char b[48]; const int n = snprintf(b, sizeof b, "%g", (double)v); s.append(b, (size_t)n); }     // ostream default: %g, 6 significant digits
  void cigar(const uint32_t *cig, int n) { static const char ops[] = "MIDNSHP=X"; for (int k = 0; k < n; k++) { u(cig[k] >> 4); s.push_back(ops[cig[k] & 15u]); } }
};

// ---- alignment strings of a record (Alignment::CreateAlignmentStrings, Alignment.h:247-333) rebuilt from its CIGAR: '=' / 'X' columns pair a read base with a
// reference base, 'I' a read base with '-', 'D' '-' with a reference base -- the column order CreateAlignmentStrings emits and AlignStringsToCigar run-length encodes
static void mp_alignment_strings(const lra_b200_record &x, const uint32_t *cig, const char *rd, const char *contig, std::string &qs, std::string &as, std::string &ts) {
  auto sm = [](unsigned char c) { switch (c) { case 'C': case 'c': case 1: case 5: return 1; case 'G': case 'g': case 2: case 6: return 2; case 'T': case 't': case 3: case 7: return 3; default: return 0; } };   // seqMap (SeqUtils.h:4-37)
  qs.clear(); as.clear(); ts.clear();
  uint32_t q = x.qStart, t = x.tStart;
  for (int k = 0; k < x.n_cigar; k++) {
    const uint32_t len = cig[k] >> 4, op = cig[k] & 15u;
    for (uint32_t l = 0; l < len; l++) {
      if (op == 1) { qs.push_back(rd[q++]); ts.push_back('-'); as.push_back(' '); }
      else if (op == 2) { qs.push_back('-'); ts.push_back(contig[t++]); as.push_back(' '); }
      else { const char a = rd[q++], b = contig[t++]; qs.push_back(a); ts.push_back(b); as.push_back(sm((unsigned char)a) != sm((unsigned char)b) ? '*' : '|'); }
    }
  }
}

// Alignment::AlignmentStringsToMD (Alignment.h:204-245), literally (including the `a and b or c` precedence of its first scan)
static void mp_md_string(const std::string &queryStr, const std::string &textStr, SamOut &o) {
  std::string query(queryStr), text(textStr);
  for (auto &c : query) c = (char)toupper((unsigned char)c);
  for (auto &c : text) c = (char)toupper((unsigned char)c);
  const size_t n = text.size();
  auto T = [&](size_t i) { return i < n ? text[i] : '\0'; };      // std::string::operator[] at size() is the terminator
  auto Q = [&](size_t i) { return i < query.size() ? query[i] : '\0'; };
  size_t s = 0;
  while (s < n) {
    size_t i = s;
    int match = 0;
    while ((i < n && T(i) == Q(i)) || T(i) == '-') { if (T(i) == Q(i)) match++; i++; if (i > n) break; }
    o.i(match);
    s = i;
    if (T(i) != Q(i) && T(i) != '-' && Q(i) != '-') { i++; if (s < n) o.put(text[s]); }
    else if (T(i) != '-' && Q(i) == '-') {
      while (i < n && T(i) != '-' && Q(i) == '-') i++;
      o.put('^'); o.put(text.data() + s, i - s);
    }
    while (i < n && T(i) == '-' && Q(i) == '=') i++;
    s = i;
  }
}

// Alignment::PrintPairwise (Alignment.h:564-589)
static void mp_print_pairwise(SamOut &o, const char *name, const char *chrom, const lra_b200_record &x, const std::string &qs, const std::string &as, const std::string &ts) {
  o.put(name); o.put('\n');
  if (x.n_blocks > 0) { o.put("Interval:\t"); o.put(chrom); o.put(':'); o.u(x.tStart); o.put('-'); o.u(x.tStart + x.tEnd); o.put('\n'); }   // refLen is the end position (refStart = 0, :256-332)
  auto w10 = [&](long long v) { char b[32]; const int n = snprintf(b, sizeof b, "%10lld", v); o.put(b, (size_t)n); };
  size_t i = 0; long long q = 0, t = 0;
  while (i < qs.size()) {
    const size_t end = qs.size() < i + 50 ? qs.size() : i + 50;
    size_t gq = 0, gt = 0;
    for (size_t k = i; k < end; k++) { gq += qs[k] == '-'; gt += ts[k] == '-'; }
    w10(q + (long long)(x.n_blocks > 0 ? x.qStart : 0)); o.put(" q: "); o.put(qs.data() + i, end - i); o.put('\n');
    q += (long long)(end - i - gq);
    o.put("              "); o.put(as.data() + i, end - i); o.put('\n');
    w10(t + (long long)(x.n_blocks > 0 ? x.tStart : 0)); o.put(" t: "); o.put(ts.data() + i, end - i); o.put('\n');
    t += (long long)(end - i - gt);
    o.put('\n');
    i = end;
  }
}

// fmt: 's' SAM (Alignment::PrintSAM, Alignment.h:658-808), 'p' PAF, 'c' PAF with CG:z: (-p pc; PrintPAF :600-656), 'b' BED (PrintBed :591-598), 'a' pairwise (PrintPairwise :564-589; needs the genome)
static void mp_format_read(SamOut &o, const lra_b200_map_opts *opts, const lra_b200_map_result *res, int r, const char *name, const char *seq, uint32_t L,
                           const std::vector<const char *> &cname, const uint64_t *contig_len, int runtime, std::string &rc, char fmt, const char *qual,
                           const char *genome = nullptr, const uint64_t *contig_off = nullptr, bool print_md = false) {
  const unsigned char *RC = mp_revcomp_table();
  const int na = res->status[r] == 0 ? res->n_aln[r] : 0;
  bool printed = false;
  if (na > 0 && res->aln_nseg[4 * r + res->aln_rank[4 * r]] > 0) {
    bool have_rc = false;
    const int lim = na < opts->PrintNumAln ? na : opts->PrintNumAln;
    for (int a = 0; a < lim; a++) {
      const int slot = res->aln_rank[4 * r + a];
      const int ns = res->aln_nseg[4 * r + slot], s0 = res->aln_seg0[4 * r + slot];
      for (int sgi = ns - 1; sgi >= 0; sgi--) {
        const lra_b200_record &x = res->records[s0 + sgi];
        printed = true;
        if (fmt == 'b') {
          o.put(x.n_blocks ? cname[x.chrom] : ""); o.put('\t'); o.u(x.tStart); o.put('\t'); o.u(x.tEnd); o.put('\t'); o.i((int)(unsigned char)x.mapq); o.put('\t'); o.put(name); o.put('\t');
          o.u(L); o.put('\t'); o.u(x.qStart); o.put('\t'); o.u(x.qEnd); o.put('\t'); o.i(x.nm); o.put('\t'); o.i(x.nmm); o.put('\t'); o.i(x.nins); o.put('\t'); o.i(x.ndel); o.put('\t');
          o.f(x.value); o.put('\t'); o.u(x.flag); o.put('\t'); o.i(x.NumOfAnchors1); o.put('\t'); o.f((float)x.NumOfAnchors1 / (float)L); o.put('\n');
          continue;
        }
        if (fmt == 'p' || fmt == 'c') {
          o.put(name); o.put('\t'); o.u(L); o.put('\t');
          if (x.strand == 0) { o.u(x.qStart); o.put('\t'); o.u(x.qEnd); o.put('\t'); }
          else { o.u(L - x.qEnd); o.put('\t'); o.u(L - x.qStart); o.put('\t'); }
          o.put(x.strand == 1 ? '-' : '+'); o.put('\t'); o.put(x.n_blocks ? cname[x.chrom] : ""); o.put('\t'); o.u(x.n_blocks && contig_len ? contig_len[x.chrom] : 0ull); o.put('\t');
          o.u(x.tStart); o.put('\t'); o.u(x.tEnd); o.put('\t'); o.i(x.nm); o.put('\t'); o.i(x.nm + x.nmm + x.ndel + x.nins); o.put('\t'); o.i((int)(unsigned char)x.mapq);
          o.put("\tOR:i:"); o.i(x.order); o.put("\tNM:i:"); o.i(x.nmm + x.ndel + x.nins); o.put("\tNX:i:"); o.i(x.nmm); o.put("\tND:i:"); o.i(x.ndel); o.put("\tTD:i:"); o.i(x.tdel);
          o.put("\tNI:i:"); o.i(x.nins); o.put("\tTI:i:"); o.i(x.tins);
          o.put("\tSD:i:"); o.i(x.nSmallDel); o.put("\tME:i:"); o.i(x.nMedDel); o.put("\tLD:i:"); o.i(x.nLargeDel); o.put("\tSI:i:"); o.i(x.nSmallIns); o.put("\tMI:i:"); o.i(x.nMedIns);
          o.put("\tLI:i:"); o.i(x.nLargeIns); o.put("\tN0:i:"); o.i(x.NumOfAnchors0); o.put("\tNV:f:"); o.f(x.value); o.put("\tAS:i:"); o.i((int)x.value);
          o.put("\tTP:A:"); o.put(x.typeofaln == 0 ? 'P' : (x.typeofaln == 1 ? 'S' : 'I'));
          if (x.NumOfAnchors1 > 0) { o.put("\tNA:i:"); o.i(x.NumOfAnchors1); }
          if (runtime > 0) { o.put("\tRT:i:"); o.i(runtime); }
          if (fmt == 'c') {
            o.put("\tCG:z:");
            if (x.preClip > 0) { o.i(x.preClip); o.put('S'); }
            o.cigar(res->cigar + x.cigar_off, x.n_cigar);
            if (x.sufClip > 0) { o.i(x.sufClip); o.put('S'); }
          }
          o.put('\n');
          continue;
        }
        const char *rd = seq;
        if (x.strand == 1) {
          if (!have_rc) { rc.resize(L); for (uint32_t k = 0; k < L; k++) rc[L - 1 - k] = (char)RC[(unsigned char)seq[k]]; have_rc = true; }
          rd = rc.data();
        }
        if (fmt == 'a') {
          std::string qs, as, ts;
          if (x.n_blocks > 0) mp_alignment_strings(x, res->cigar + x.cigar_off, rd, genome + contig_off[x.chrom], qs, as, ts);
          mp_print_pairwise(o, name, x.n_blocks ? cname[x.chrom] : "", x, qs, as, ts);
          continue;
        }
        o.put(name); o.put('\t');
        if (x.n_blocks == 0) { o.put("4\t*\t0\t0\t*\t*\t0\t0\t"); o.put(rd, L); o.put('\t'); if (qual) o.put(qual, L); else o.put('*'); }
        else {
          o.u(x.flag); o.put('\t'); o.put(cname[x.chrom]); o.put('\t'); o.u(x.tStart + 1u); o.put('\t'); o.u((unsigned char)x.mapq); o.put('\t');
          char clipOp = 'S';
          if (x.supplementary && opts->hardClip) clipOp = 'H';
          if (x.preClip > 0) { o.i(x.preClip); o.put(clipOp); }
          o.cigar(res->cigar + x.cigar_off, x.n_cigar);
          if (x.sufClip > 0) { o.i(x.sufClip); o.put(clipOp); }
          o.put("\t*\t0\t"); o.u(x.tEnd - x.tStart); o.put('\t');
          if (!x.supplementary) o.put(rd, L);
          else if (opts->hardClip) o.put(rd + x.qStart, x.qEnd - x.qStart);
          else o.put(rd, L);
          // QUAL (Alignment.h:719-732): as given in the input, not reversed with the strand; hard-clipped like SEQ
          o.put('\t');
          if (!qual || qual[0] == '*') o.put('*');
          else if (x.supplementary && opts->hardClip) o.put(qual + x.qStart, x.qEnd - x.qStart);
          else o.put(qual, L);
          o.put("\tNM:i:"); o.i(x.nmm + x.ndel + x.nins); o.put("\tMM:i:"); o.i(x.nmm + x.ndel + x.nins); o.put("\tNX:i:"); o.i(x.nmm); o.put("\tND:i:"); o.i(x.ndel);
          o.put("\tTD:i:"); o.i(x.tdel); o.put("\tNI:i:"); o.i(x.nins); o.put("\tTI:i:"); o.i(x.tins); o.put("\tNV:f:"); o.f(x.value); o.put("\tAS:i:"); o.i((int)x.value);
          o.put("\tAO:i:"); o.i(x.order); o.put("\tN0:i:"); o.i(x.NumOfAnchors0); o.put("\tRT:i:"); o.i(runtime);
          o.put("\tTP:A:"); o.put(x.typeofaln == 0 ? 'P' : (x.typeofaln == 1 ? 'S' : 'I'));
          o.put("\tSD:i:"); o.i(x.nSmallDel); o.put("\tME:i:"); o.i(x.nMedDel); o.put("\tLD:i:"); o.i(x.nLargeDel); o.put("\tSI:i:"); o.i(x.nSmallIns); o.put("\tMI:i:"); o.i(x.nMedIns);
          o.put("\tLI:i:"); o.i(x.nLargeIns);
          if (print_md) {      // opts.printMD (Alignment.h:763-767)
            std::string qs, as, ts;
            mp_alignment_strings(x, res->cigar + x.cigar_off, rd, genome + contig_off[x.chrom], qs, as, ts);
            o.put("\tMD:Z:"); mp_md_string(qs, ts, o);
          }
          if (ns > 1) o.put("\tSA:Z:");
          for (int ag = ns - 1; ag >= 0; ag--) {
            if (ag == sgi) continue;
            const lra_b200_record &y = res->records[s0 + ag];
            o.put(y.n_blocks == 0 ? "*" : cname[y.chrom]); o.put(','); o.u(y.tStart + 1u); o.put(','); o.put(y.strand == 0 ? '+' : '-'); o.put(',');
            if (y.preClip > 0) { o.i(y.preClip); o.put('S'); }
            o.cigar(res->cigar + y.cigar_off, y.n_cigar);
            if (y.sufClip > 0) { o.i(y.sufClip); o.put('S'); }
            o.put(','); o.u((unsigned char)y.mapq); o.put(','); o.i(y.nm); o.put(';');
          }
        }
        o.put('\n');
      }
    }
  }
  if (!printed && fmt == 's') {      // output_unaligned -> SimplePrintSAM of an Alignment without blocks (SAM only, Mapping_ultility.h:457-463)
    o.put(name); o.put("\t4\t*\t0\t0\t*\t*\t0\t0\t"); o.put(seq, L); o.put('\t'); if (qual) o.put(qual, L); else o.put('*'); o.put('\n');
  }
}

extern "C" int64_t lra_b200_format_records(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                           const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len, int32_t n_contigs,
                                           int32_t fmt, int32_t runtime, char *out, int64_t cap);
extern "C" int64_t lra_b200_format_sam(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                       const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, int32_t n_contigs, int32_t runtime, char *out,
                                       int64_t cap) {
  return lra_b200_format_records(opts, res, n_reads, names, reads_ascii, read_off, read_len, contig_names, nullptr, n_contigs, 's', runtime, out, cap);
}

extern "C" int64_t lra_b200_format_records_qual(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                                const char *quals_ascii, const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len,
                                                int32_t n_contigs, int32_t fmt, int32_t runtime, char *out, int64_t cap);
extern "C" int64_t lra_b200_format_records(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                           const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len, int32_t n_contigs,
                                           int32_t fmt, int32_t runtime, char *out, int64_t cap) {
  return lra_b200_format_records_qual(opts, res, n_reads, names, reads_ascii, nullptr, read_off, read_len, contig_names, contig_len, n_contigs, fmt, runtime, out, cap);
}
extern "C" int64_t lra_b200_format_records_ref(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                               const char *quals_ascii, const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len,
                                               int32_t n_contigs, const char *genome_ascii, const uint64_t *contig_off, int32_t fmt, int32_t print_md, int32_t runtime, char *out,
                                               int64_t cap);
extern "C" int64_t lra_b200_format_records_qual(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                                const char *quals_ascii, const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len,
                                                int32_t n_contigs, int32_t fmt, int32_t runtime, char *out, int64_t cap) {
  if (fmt == 'a') return 0;      // the pairwise view needs the reference bases: lra_b200_format_records_ref
  return lra_b200_format_records_ref(opts, res, n_reads, names, reads_ascii, quals_ascii, read_off, read_len, contig_names, contig_len, n_contigs, nullptr, nullptr, fmt, 0, runtime, out, cap);
}
extern "C" int64_t lra_b200_format_records_ref(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                               const char *quals_ascii, const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len,
                                               int32_t n_contigs, const char *genome_ascii, const uint64_t *contig_off, int32_t fmt, int32_t print_md, int32_t runtime, char *out,
                                               int64_t cap) {
  if (fmt != 's' && fmt != 'p' && fmt != 'c' && fmt != 'b' && fmt != 'a') return 0;
  if ((fmt == 'a' || print_md) && (!genome_ascii || !contig_off)) return 0;
  if (!opts || !res || n_reads < 0 || !names || !reads_ascii || !read_off || !read_len || !contig_names) return 0;
  std::vector<const char *> cname(n_contigs);
  { const char *p = contig_names; for (int c = 0; c < n_contigs; c++) { cname[c] = p; p += strlen(p) + 1; } }
  std::vector<const char *> rname(n_reads);
  { const char *p = names; for (int r = 0; r < n_reads; r++) { rname[r] = p; p += strlen(p) + 1; } }
  // host threads over contiguous, base-balanced runs of reads (the text of a read depends on that read only); pieces are concatenated in input order
  unsigned long long total = 0; for (int r = 0; r < n_reads; r++) total += read_len[r];
  int T = (int)std::thread::hardware_concurrency(); if (T < 1) T = 1; if (T > 32) T = 32;
  if (total < (1u << 20)) T = 1;
  if (getenv("LRA_B200_SAM_THREADS")) T = atoi(getenv("LRA_B200_SAM_THREADS"));
  if (T < 1) T = 1;
  if (T > n_reads) T = n_reads > 0 ? n_reads : 1;
  std::vector<int> cut(T + 1, n_reads); cut[0] = 0;
  { unsigned long long acc = 0; int t = 1; for (int r = 0; r < n_reads && t < T; r++) { acc += read_len[r]; if (acc * T >= total * t) cut[t++] = r + 1; } }
  std::vector<SamOut> piece(T);
  auto work = [&](int t) {
    std::string rc;
    unsigned long long b = 0; for (int r = cut[t]; r < cut[t + 1]; r++) b += read_len[r];
    piece[t].s.reserve((size_t)(b + b / 2) + 4096);
    for (int r = cut[t]; r < cut[t + 1]; r++) mp_format_read(piece[t], opts, res, r, rname[r], reads_ascii + read_off[r], read_len[r], cname, contig_len, runtime, rc, (char)fmt, quals_ascii ? quals_ascii + read_off[r] : nullptr,
                                                                     genome_ascii, contig_off, print_md != 0 && fmt == 's');
  };
  if (T == 1) work(0);
  else { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(work, t); for (auto &x : th) x.join(); }
  size_t n = 0; for (auto &p : piece) n += p.s.size();
  if ((int64_t)n > cap || !out) return -(int64_t)n;
  std::vector<size_t> at(T + 1, 0);
  for (int t = 0; t < T; t++) at[t + 1] = at[t] + piece[t].s.size();
  auto copy = [&](int t) { memcpy(out + at[t], piece[t].s.data(), piece[t].s.size()); };
  if (T == 1) copy(0);
  else { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(copy, t); for (auto &x : th) x.join(); }
  return (int64_t)n;
}

struct lra_b200_readset;
extern "C" void lra_b200_readset_free(lra_b200_ctx *ctx, lra_b200_readset *rs);
// ---- the mapper ---------------------------------------------------------------------------------------------------------------------------------
struct lra_b200_mapper {
  lra_b200_map_opts opts;
  lra_b200_seq *genome = nullptr;
  lra_b200_index *index = nullptr;
  lra_b200_lindex *gl = nullptr;
  unsigned long long *d_hdr = nullptr; int n_hdr = 0;
  std::vector<uint64_t> h_hdr;
  lra::mp::Pwl *d_pwl = nullptr;
  float log_lut[2001];
  float logf_len[8];
  // per batch
  struct lra_b200_readset *own = nullptr; // the batch of lra_b200_map_batch (host buffers in, records out)
  lra_b200_lindex *rl[2] = {nullptr, nullptr};
  int last_reads = 0, last_S = 0, last_kerr = 0; uint64_t last_ncig = 0;
  DevBuf b[40];
  int *h_pin = nullptr;
};

struct lra_b200_index_internal_view { const unsigned long long *t; const uint32_t *pos; uint64_t n; };

namespace lra { namespace mp {
// SegRec -> the a19 / a21 segment arrays
__global__ void seg_to_ir_kernel(const SegRec *seg, int n, const unsigned long long *read_off, const uint32_t *read_len, unsigned long long rc_shift,
                                 const unsigned long long *hdr_pos, unsigned long long *blk_off, int32_t *blk_cnt, uint32_t *q_base, uint32_t *t_base, int32_t *rlen,
                                 int32_t *clen) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const SegRec x = seg[s];
  blk_off[s] = x.blk_off; blk_cnt[s] = x.blk_cnt;
  q_base[s] = (uint32_t)(read_off[x.read] + (x.strand ? rc_shift : 0ull));
  t_base[s] = (uint32_t)hdr_pos[x.chrom];
  rlen[s] = (int32_t)read_len[x.read]; clen[s] = (int32_t)(hdr_pos[x.chrom + 1] - hdr_pos[x.chrom]);
}
}}

extern "C" void lra_b200_mapper_destroy(lra_b200_ctx *ctx, lra_b200_mapper *m) {
  if (!m) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (m->genome) lra_b200_seq_free(ctx, m->genome);
  if (m->index) lra_b200_index_free(ctx, m->index);
  if (m->gl) lra_b200_lindex_free(ctx, m->gl);
  if (m->own) lra_b200_readset_free(ctx, m->own);
  for (int s = 0; s < 2; s++) if (m->rl[s]) lra_b200_lindex_free(ctx, m->rl[s]);
  if (m->d_hdr) cudaFree(m->d_hdr);
  if (m->d_pwl) cudaFree(m->d_pwl);
  for (DevBuf &b : m->b) if (b.p) cudaFree(b.p);
  if (m->h_pin) cudaFreeHost(m->h_pin);
  delete m;
}

extern "C" int lra_b200_mapper_create(lra_b200_ctx *ctx, const lra_b200_map_opts *opts, const char *genome_ascii, uint64_t genome_len, const uint64_t *hdr_pos,
                                      int32_t n_contigs, const uint64_t *mms_t, const uint32_t *mms_pos, uint64_t n_mms, const uint64_t *gli_seq_offsets,
                                      const uint64_t *gli_tuple_boundaries, int32_t gli_n_regions, const uint32_t *gli_minimizers, uint64_t gli_n_min,
                                      lra_b200_mapper **out) {
  if (!ctx || !opts || !genome_ascii || !hdr_pos || n_contigs < 1 || !mms_t || !mms_pos || !out) return fail(ctx, LRA_B200_EINVAL, "mapper_create: NULL argument");
  if (!opts->bypassClustering && !opts->HighlyAccurate)
    return fail(ctx, LRA_B200_EINVAL, "mapper_create: neither bypassClustering (-ONT, -CLR: MapRead_lowacc) nor HighlyAccurate (-CCS, -CONTIG: MapRead_highacc) is set");
  if (hdr_pos[n_contigs] != genome_len || genome_len >= (1ull << 32)) return fail(ctx, LRA_B200_EINVAL, "mapper_create: header offsets do not cover the genome (or >= 2^32 bases)");
  *out = nullptr;
  CU(cudaSetDevice(ctx->device));
  {   // the packed alphabet is A, C, G, T + one "N" code: IUPAC ambiguity bytes compare equal to N here but as raw bytes in the reference (Checkbp, RefineSpace, strncmp)
    bool plain[256]; for (int i = 0; i < 256; i++) plain[i] = false;
    for (const char *c = "ACGTNacgtn"; *c; c++) plain[(unsigned char)*c] = true;
    unsigned long long other = 0;
    for (uint64_t i = 0; i < genome_len; i++) other += plain[(unsigned char)genome_ascii[i]] ? 0 : 1;
    if (other) fprintf(stderr, "lra_b200: warning: %llu reference bases outside ACGTN are treated as N; alignments that overlap them may differ from the reference's\n", other);
  }
  lra_b200_mapper *m = new lra_b200_mapper();
  m->opts = *opts;
  int rc;
  if ((rc = lra_b200_seq_upload(ctx, genome_ascii, genome_len, &m->genome))) { lra_b200_mapper_destroy(ctx, m); return rc; }
  if ((rc = lra_b200_index_upload(ctx, mms_t, mms_pos, n_mms, &m->index))) { lra_b200_mapper_destroy(ctx, m); return rc; }
  std::vector<uint64_t> starts(n_contigs); std::vector<uint32_t> lens(n_contigs);
  for (int c = 0; c < n_contigs; c++) { starts[c] = hdr_pos[c]; lens[c] = (uint32_t)(hdr_pos[c + 1] - hdr_pos[c]); }
  if (gli_seq_offsets && gli_tuple_boundaries && gli_n_regions > 0) {
    (void)gli_n_min;
    rc = lra_b200_lindex_upload(ctx, starts.data(), lens.data(), n_contigs, opts->localIndexWindow, gli_seq_offsets, gli_tuple_boundaries, gli_minimizers,
                                (uint64_t)gli_n_regions - 1, &m->gl);
  } else rc = lra_b200_lindex_build(ctx, m->genome, starts.data(), lens.data(), n_contigs, opts->smallK, opts->smallW, opts->localIndexWindow, opts->localIndexMaxFreq, &m->gl);
  if (rc) { lra_b200_mapper_destroy(ctx, m); return rc; }
  m->n_hdr = n_contigs; m->h_hdr.assign(hdr_pos, hdr_pos + n_contigs + 1);
  if (cudaMalloc(&m->d_hdr, (size_t)(n_contigs + 1) * 8) != cudaSuccess || cudaMalloc(&m->d_pwl, sizeof(lra::mp::Pwl)) != cudaSuccess ||
      cudaHostAlloc((void **)&m->h_pin, 256, cudaHostAllocDefault) != cudaSuccess) { lra_b200_mapper_destroy(ctx, m); return fail(ctx, LRA_B200_ECUDA, "mapper_create: allocation failed"); }
  lra::mp::Pwl hp; int64_t stops[25]; float slope[25], inter[25];
  lra_b200_init_pwl(opts->gapopen, opts->gapextend, opts->gaproot, opts->gapCeiling1, opts->gapCeiling2, stops, slope, inter);
  mp_fill_pwl(hp, stops, slope, inter, opts->gapCeiling1, opts->gapCeiling2);
  cudaMemcpyAsync(m->d_hdr, hdr_pos, (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(m->d_pwl, &hp, sizeof hp, cudaMemcpyHostToDevice, ctx->stream);
  CU(cudaStreamSynchronize(ctx->stream));
  // CreateLookUpTable (LogLookUpTable.h:9-15) with the host logf
  { int k = 0; for (int i = 1; i <= 10001; i += 5) m->log_lut[k++] = logf((float)i); }
  for (int i = 0; i < 8; i++) m->logf_len[i] = i > 0 ? logf((float)i) : 0.0f;
  *out = m;
  return LRA_B200_OK;
}


// ---- a batch of reads resident on the device (packed forward strands + reverse complements, descriptors, longest-first order) ----------------------
struct lra_b200_readset {
  lra_b200_seq *reads = nullptr;          // forward strands at [0, Npad), reverse complements at [Npad, 2 Npad)
  DevBuf off, len, order;
  std::vector<uint64_t> h_off, h_rc_off; std::vector<uint32_t> h_len;
  int n_reads = 0; unsigned long long Npad = 0, total_bases = 0; uint32_t maxL = 0;
};

extern "C" void lra_b200_readset_free(lra_b200_ctx *ctx, lra_b200_readset *rs) {
  if (!rs) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (rs->reads) lra_b200_seq_free(ctx, rs->reads);
  for (DevBuf *b : {&rs->off, &rs->len, &rs->order}) if (b->p) cudaFree(b->p);
  delete rs;
}

static int readset_fill(lra_b200_ctx *ctx, lra_b200_readset *rs, const char *reads_ascii, uint64_t reads_len, const uint64_t *read_off, const uint32_t *read_len, int32_t n_reads) {
  CU(cudaSetDevice(ctx->device));
  rs->n_reads = n_reads; rs->maxL = 0; rs->total_bases = 0; rs->Npad = 0;
  if (n_reads == 0) return LRA_B200_OK;
  uint32_t maxL = 0; unsigned long long total_bases = 0;
  for (int r = 0; r < n_reads; r++) {
    if (read_off[r] + read_len[r] > reads_len) return fail(ctx, LRA_B200_EINVAL, "readset_upload: read %d ends beyond the buffer", r);
    if (r && read_off[r] < read_off[r - 1] + read_len[r - 1]) return fail(ctx, LRA_B200_EINVAL, "readset_upload: reads overlap or are not in ascending order at %d", r);
    maxL = read_len[r] > maxL ? read_len[r] : maxL; total_bases += read_len[r];
  }
  const unsigned long long Npad = ((unsigned long long)reads_len + 127ull) & ~63ull;
  if (2 * Npad >= (1ull << 32)) return fail(ctx, LRA_B200_EINVAL, "readset_upload: more than 2^31 bases in one batch");
  cudaStream_t st = ctx->stream;
  int rc;
  // ---- reads: pack the forward strands into [0, Npad), reverse complements into [Npad, 2 Npad)
  if (!rs->reads) rs->reads = new lra_b200_seq();
  if ((rc = seq_reserve(ctx, rs->reads, 2 * Npad))) return rc;
  CU(cudaMemsetAsync(rs->reads->b2, 0, (rs->reads->cap_groups * 2 + 8) * 4, st));
  CU(cudaMemsetAsync(rs->reads->nm, 0xFF, (rs->reads->cap_groups + 8) * 4, st));
  if (reads_len + 64 > rs->reads->ascii_cap) {
    if (rs->reads->ascii_dev) { CU(cudaStreamSynchronize(st)); CU(cudaFree(rs->reads->ascii_dev)); rs->reads->ascii_dev = nullptr; }
    const uint64_t cap = reads_len + reads_len / 4 + 256;
    CU(cudaMalloc((void **)&rs->reads->ascii_dev, cap)); rs->reads->ascii_cap = cap;
  }
  CU(cudaMemcpyAsync(rs->reads->ascii_dev, reads_ascii, reads_len, cudaMemcpyDefault, st));   // host or device source (UVA)
  { const uint64_t groups = (reads_len + 31) / 32 + 1;
    lra::seq_pack_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(rs->reads->ascii_dev, reads_len, rs->reads->b2, rs->reads->nm, groups); ctx->launches++; }
  rs->reads->n = 2 * Npad;
  if ((rc = ensure(ctx, rs->off, (size_t)n_reads * 8)) || (rc = ensure(ctx, rs->len, (size_t)n_reads * 4))) return rc;
  CU(cudaMemcpyAsync(rs->off.p, read_off, (size_t)n_reads * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(rs->len.p, read_len, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
  lra::seq_revcomp_kernel<<<(unsigned)((n_reads + 7) / 8), 256, 0, st>>>(lra::SeqView{rs->reads->b2, rs->reads->nm, Npad}, (const unsigned long long *)rs->off.p,
                                                                         (const uint32_t *)rs->len.p, n_reads, rs->reads->b2 + Npad / 16, rs->reads->nm + Npad / 32);
  ctx->launches++;
  CU(cudaGetLastError());
  // longest reads first (the tail of the batch is then made of short reads)
  if ((rc = ensure(ctx, rs->order, (size_t)n_reads * 4))) return rc;
  { std::vector<int> ord(n_reads); for (int r = 0; r < n_reads; r++) ord[r] = r;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b2) { return read_len[a] > read_len[b2]; });
    CU(cudaMemcpyAsync(rs->order.p, ord.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, st)); CU(cudaStreamSynchronize(st)); }
  rs->h_off.assign(read_off, read_off + n_reads); rs->h_len.assign(read_len, read_len + n_reads); rs->h_rc_off.resize(n_reads);
  for (int r = 0; r < n_reads; r++) rs->h_rc_off[r] = read_off[r] + Npad;
  rs->Npad = Npad; rs->maxL = maxL; rs->total_bases = total_bases;
  return LRA_B200_OK;
}

extern "C" int lra_b200_readset_upload(lra_b200_ctx *ctx, const char *reads_ascii, uint64_t reads_len, const uint64_t *read_off, const uint32_t *read_len, int32_t n_reads,
                                       lra_b200_readset **out) {
  if (!ctx || !out || n_reads < 0 || (n_reads && (!reads_ascii || !read_off || !read_len))) return fail(ctx, LRA_B200_EINVAL, "readset_upload: NULL argument");
  *out = nullptr;
  lra_b200_readset *rs = new lra_b200_readset();
  const int rc = readset_fill(ctx, rs, reads_ascii, reads_len, read_off, read_len, n_reads);
  if (rc) { lra_b200_readset_free(ctx, rs); return rc; }
  *out = rs;
  return LRA_B200_OK;
}

// every kernel of the path over a resident batch; the records stay in the mapper's device buffers until lra_b200_map_download
extern "C" int lra_b200_map_resident(lra_b200_ctx *ctx, lra_b200_mapper *m, const lra_b200_readset *rs) {
  using namespace lra::mp;
  if (!ctx || !m || !rs) return fail(ctx, LRA_B200_EINVAL, "map_resident: NULL argument");
  CU(cudaSetDevice(ctx->device));
  const int n_reads = rs->n_reads; const unsigned long long Npad = rs->Npad, total_bases = rs->total_bases; const uint32_t maxL = rs->maxL;
  m->last_reads = n_reads; m->last_S = 0; m->last_ncig = 0; m->last_kerr = 0;
  std::vector<lra_b200_kernel_stat> all;
  if (n_reads == 0) { ctx->stats.clear(); return LRA_B200_OK; }
  cudaStream_t st = ctx->stream;
  int rc;
  DevBuf *B = m->b;
  // ---- a12: LocalIndex::IndexSeq of every read, both strands (Map_lowacc.h:246-250).  MapRead_highacc builds them too (Map_highacc.h:399-403) but reads them only in
  // REFINEclusters, i.e. for the few reads with a sparse cluster: those reads come back with MP_NEED_LIDX, are indexed, and mapped again below.
  ctx->keep_stats = false;
  const bool lazy_lidx = m->opts.HighlyAccurate != 0 && !getenv("LRA_B200_MAP_EAGER_LIDX");
  if (!lazy_lidx) {
    if ((rc = lra_b200_lindex_build(ctx, rs->reads, rs->h_off.data(), rs->h_len.data(), n_reads, m->opts.smallK, m->opts.smallW, m->opts.localIndexWindow, m->opts.localIndexMaxFreq, &m->rl[0]))) return rc;
    all.insert(all.end(), ctx->stats.begin(), ctx->stats.end());
    if ((rc = lra_b200_lindex_build(ctx, rs->reads, rs->h_rc_off.data(), rs->h_len.data(), n_reads, m->opts.smallK, m->opts.smallW, m->opts.localIndexWindow, m->opts.localIndexMaxFreq, &m->rl[1]))) return rc;
    all.insert(all.end(), ctx->stats.begin(), ctx->stats.end());
  }
  // ---- the mapper worker kernel
  const size_t seg_cap = (size_t)n_reads * 3 + 1024 + (m->opts.readType == 3 ? 65536 : 0);      // contigs: many segments per read
  const size_t blk_cap = (size_t)(total_bases / 2) + (size_t)n_reads * 64 + 4096;
  // one CTA of `bw` warps per SM (phase-aligned groups of reads, mp_phase); LRA_B200_MAP_BLOCK_WARPS / _BLOCKS_PER_SM for experiments
  int bw = MP_BLOCK_THREADS / 32; if (getenv("LRA_B200_MAP_BLOCK_WARPS")) bw = atoi(getenv("LRA_B200_MAP_BLOCK_WARPS"));
  if (bw < 1) bw = 1; if (bw > MP_BLOCK_THREADS / 32) bw = MP_BLOCK_THREADS / 32;
  int bps = 1; if (getenv("LRA_B200_MAP_BLOCKS_PER_SM")) bps = atoi(getenv("LRA_B200_MAP_BLOCKS_PER_SM"));
  if (bps < 1) bps = 1; if (bps * bw > MP_BLOCK_THREADS / 32) bps = (MP_BLOCK_THREADS / 32) / bw;
  int blocks = ctx->n_sm * bps;
  if ((long long)blocks * bw > (long long)n_reads) blocks = (n_reads + bw - 1) / bw;
  int warps = blocks * bw;
  // worker scratch, measured on ONT / CLR reads: peak 14.4 MB for a 100 kb read, ~150 B per base (SparseDP sub-problems dominate).  The first pass
  // gives every warp 4 MB + 256 B per base of the longest read (measured peak: 21 MB for a 100 kb read); a read that still runs out (status MP_ERR_ARENA) is mapped again below with 8x
  // that on fewer warps.  The arena is kept across batches.
  size_t per = (size_t)maxL * 256 + (4u << 20);
  if (getenv("LRA_B200_MAP_ARENA_MB")) per = (size_t)atoi(getenv("LRA_B200_MAP_ARENA_MB")) << 20;
  size_t free_b = 0, tot_b = 0; cudaMemGetInfo(&free_b, &tot_b);
  // the worker arenas take at most 70 % of what is free (the rest of the batch -- records, blocks, a19 / a21 buffers -- needs a few GB); the
  // allocation is exact (no growth slack) and, should it still fail (another allocator holding memory), retried with fewer CTAs
  const size_t budget = (free_b + B[9].cap) / 10 * 7;
  while ((size_t)warps * per > budget && blocks > 1) { blocks--; warps = blocks * bw; }
  if (B[9].cap < (size_t)warps * per) {
    if (B[9].p) { CU(cudaStreamSynchronize(st)); CU(cudaFree(B[9].p)); B[9].p = nullptr; B[9].cap = 0; }
    for (;;) {
      if (cudaMalloc(&B[9].p, (size_t)warps * per) == cudaSuccess) { B[9].cap = (size_t)warps * per; break; }
      cudaGetLastError();
      B[9].p = nullptr;
      if (blocks <= 1) return fail(ctx, LRA_B200_ECUDA, "map_batch: no memory for the worker arenas (%zu bytes per warp)", per);
      blocks -= (blocks + 3) / 4; warps = blocks * bw;
    }
  }
  if ((rc = ensure(ctx, B[2], (size_t)n_reads * 4)) || (rc = ensure(ctx, B[3], (size_t)n_reads * 4)) || (rc = ensure(ctx, B[4], (size_t)n_reads * 16)) ||
      (rc = ensure(ctx, B[5], (size_t)n_reads * 16)) || (rc = ensure(ctx, B[6], seg_cap * sizeof(SegRec))) || (rc = ensure(ctx, B[7], blk_cap * 12)) ||
      (rc = ensure(ctx, B[8], 256)) ||
      (rc = ensure(ctx, B[27], (size_t)(warps + 4) * lra::mp::kProfStages * 8)))
    return rc;
  CU(cudaMemsetAsync(B[8].p, 0, 256, st));
  CU(cudaMemsetAsync(B[27].p, 0, (size_t)(warps + 4) * lra::mp::kProfStages * 8, st));
  MapBatch mb;
  mb.C.o = m->opts; mb.C.pwl = m->d_pwl; mb.C.prof = getenv("LRA_B200_MAP_PROFILE") ? (unsigned long long *)B[27].p : nullptr;
  mb.C.ix.genome = lra::SeqView{m->genome->b2, m->genome->nm, m->genome->n}; mb.C.ix.hdr_pos = m->d_hdr; mb.C.ix.n_hdr = m->n_hdr;
  mb.C.ix.idx_t = (const unsigned long long *)m->index->t; mb.C.ix.idx_pos = m->index->pos; mb.C.ix.n_idx = (long long)m->index->n;
  mb.C.ix.gl = lidx_view(m->gl);
  mb.C.rd.fwd = lra::SeqView{rs->reads->b2, rs->reads->nm, Npad}; mb.C.rd.rc = lra::SeqView{rs->reads->b2 + Npad / 16, rs->reads->nm + Npad / 32, Npad};
  mb.C.rd.read_off = (const unsigned long long *)rs->off.p; mb.C.rd.read_len = (const uint32_t *)rs->len.p; mb.C.rd.n_reads = n_reads;
  mb.C.rd.lidx_slot = nullptr;
  if (lazy_lidx) { memset(&mb.C.rd.rd[0], 0, sizeof mb.C.rd.rd[0]); memset(&mb.C.rd.rd[1], 0, sizeof mb.C.rd.rd[1]); }
  else { mb.C.rd.rd[0] = lidx_view(m->rl[0]); mb.C.rd.rd[1] = lidx_view(m->rl[1]); }
  // (the reverse-complement image was built over arena offsets read_off + Npad; the worker only uses window offsets relative to the image's own seq_start)
  mb.out.status = (int *)B[2].p; mb.out.n_chains = (int *)B[3].p; mb.out.chain_nseg = (int *)B[4].p; mb.out.chain_seg0 = (int *)B[5].p;
  mb.out.seg = (SegRec *)B[6].p; mb.out.seg_cap = (int)seg_cap; mb.out.seg_cursor = (unsigned long long *)B[8].p; mb.out.blocks = (uint32_t *)B[7].p; mb.out.blk_cap = blk_cap;
  mb.out.blk_cursor = (unsigned long long *)((char *)B[8].p + 8); mb.out.err = (int *)((char *)B[8].p + 16); mb.out.peak = (unsigned long long *)((char *)B[8].p + 24);
  mb.phase_mask = 0xffffffffu; if (getenv("LRA_B200_MAP_PHASE_MASK")) mb.phase_mask = (unsigned)strtoul(getenv("LRA_B200_MAP_PHASE_MASK"), nullptr, 16);
  mb.arena = (unsigned char *)B[9].p; mb.arena_per_warp = per; mb.work = (int *)((char *)B[8].p + 32); mb.order = (const int *)rs->order.p; mb.n_work = n_reads;
  cudaEventRecord(ctx->ev[0], st);
  map_reads_kernel<<<(unsigned)blocks, (unsigned)(bw * 32), 0, st>>>(mb);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  unsigned long long hcur[4];
  { // second pass for the reads whose scratch did not fit
    std::vector<int> hst(n_reads);
    CU(cudaMemcpyAsync(hst.data(), B[2].p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    { lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "map_reads"); cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]);
      s2.jobs = (uint64_t)n_reads; s2.algo_bytes = total_bases + 2 * ((total_bases + 3) / 4); all.push_back(s2); }
    if (lazy_lidx) {
      std::vector<int> need;
      for (int r = 0; r < n_reads; r++) if (hst[r] == MP_NEED_LIDX) need.push_back(r);
      if (!need.empty()) {
        const int nn = (int)need.size();
        std::vector<uint64_t> o0(nn), o1(nn); std::vector<uint32_t> ln(nn); std::vector<int> slot(n_reads, -1);
        for (int i = 0; i < nn; i++) { o0[i] = rs->h_off[need[i]]; o1[i] = rs->h_rc_off[need[i]]; ln[i] = rs->h_len[need[i]]; slot[need[i]] = i; }
        if ((rc = lra_b200_lindex_build(ctx, rs->reads, o0.data(), ln.data(), nn, m->opts.smallK, m->opts.smallW, m->opts.localIndexWindow, m->opts.localIndexMaxFreq, &m->rl[0]))) return rc;
        all.insert(all.end(), ctx->stats.begin(), ctx->stats.end());
        if ((rc = lra_b200_lindex_build(ctx, rs->reads, o1.data(), ln.data(), nn, m->opts.smallK, m->opts.smallW, m->opts.localIndexWindow, m->opts.localIndexMaxFreq, &m->rl[1]))) return rc;
        all.insert(all.end(), ctx->stats.begin(), ctx->stats.end());
        std::stable_sort(need.begin(), need.end(), [&](int a, int b2) { return rs->h_len[a] > rs->h_len[b2]; });
        if ((rc = ensure(ctx, B[31], (size_t)n_reads * 4)) || (rc = ensure(ctx, B[32], (size_t)nn * 4))) return rc;
        CU(cudaMemcpyAsync(B[31].p, slot.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(B[32].p, need.data(), (size_t)nn * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync((char *)B[8].p + 32, 0, 4, st));
        mb.C.rd.rd[0] = lidx_view(m->rl[0]); mb.C.rd.rd[1] = lidx_view(m->rl[1]); mb.C.rd.lidx_slot = (const int *)B[31].p;
        MapBatch mbs = mb; mbs.order = (const int *)B[32].p; mbs.n_work = nn;
        // few reads: spread them over all SMs (CTAs of fewer warps) instead of filling a few 24-warp CTAs
        int bw_s = (nn + ctx->n_sm - 1) / ctx->n_sm; if (bw_s < 1) bw_s = 1; if (bw_s > bw) bw_s = bw;
        int blocks_s = (nn + bw_s - 1) / bw_s; if (blocks_s > blocks) blocks_s = blocks;
        cudaEventRecord(ctx->ev[0], st);
        map_reads_kernel<<<(unsigned)blocks_s, (unsigned)(bw_s * 32), 0, st>>>(mbs);
        cudaEventRecord(ctx->ev[1], st);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hst.data(), B[2].p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        lra_b200_kernel_stat s4; memset(&s4, 0, sizeof s4); snprintf(s4.name, sizeof s4.name, "map_reads(REFINEclusters reads)"); cudaEventElapsedTime(&s4.ms, ctx->ev[0], ctx->ev[1]);
        s4.jobs = (uint64_t)nn; all.push_back(s4);
      }
    }
    std::vector<int> redo;
    for (int r = 0; r < n_reads; r++) if (hst[r] == MP_ERR_ARENA) redo.push_back(r);
    if (!redo.empty() && !getenv("LRA_B200_MAP_NO_RETRY")) {
      std::stable_sort(redo.begin(), redo.end(), [&](int a, int b2) { return rs->h_len[a] > rs->h_len[b2]; });
      const size_t per2 = per * 8;
      int blocks2 = (int)((redo.size() + bw - 1) / bw); if (blocks2 > ctx->n_sm) blocks2 = ctx->n_sm;
      while ((size_t)blocks2 * bw * per2 > B[9].cap && blocks2 > 1) blocks2--;
      int bw2 = bw; while ((size_t)blocks2 * bw2 * per2 > B[9].cap && bw2 > 1) bw2--;
      if ((size_t)blocks2 * bw2 * per2 <= B[9].cap) {
        if (lazy_lidx) {      // a read that ran out of scratch may still reach REFINEclusters: index the reads of this pass
          const int nn = (int)redo.size();
          std::vector<uint64_t> o0(nn), o1(nn); std::vector<uint32_t> ln(nn); std::vector<int> slot(n_reads, -1);
          std::vector<int> asc(redo); std::sort(asc.begin(), asc.end());
          for (int i = 0; i < nn; i++) { o0[i] = rs->h_off[asc[i]]; o1[i] = rs->h_rc_off[asc[i]]; ln[i] = rs->h_len[asc[i]]; slot[asc[i]] = i; }
          if ((rc = lra_b200_lindex_build(ctx, rs->reads, o0.data(), ln.data(), nn, m->opts.smallK, m->opts.smallW, m->opts.localIndexWindow, m->opts.localIndexMaxFreq, &m->rl[0]))) return rc;
          if ((rc = lra_b200_lindex_build(ctx, rs->reads, o1.data(), ln.data(), nn, m->opts.smallK, m->opts.smallW, m->opts.localIndexWindow, m->opts.localIndexMaxFreq, &m->rl[1]))) return rc;
          if ((rc = ensure(ctx, B[31], (size_t)n_reads * 4))) return rc;
          CU(cudaMemcpyAsync(B[31].p, slot.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
          CU(cudaStreamSynchronize(st));
          mb.C.rd.rd[0] = lidx_view(m->rl[0]); mb.C.rd.rd[1] = lidx_view(m->rl[1]); mb.C.rd.lidx_slot = (const int *)B[31].p;
        }
        if ((rc = ensure(ctx, B[28], redo.size() * 4))) return rc;
        CU(cudaMemcpyAsync(B[28].p, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync((char *)B[8].p + 32, 0, 4, st));
        MapBatch mb2 = mb; mb2.order = (const int *)B[28].p; mb2.n_work = (int)redo.size(); mb2.arena_per_warp = per2;
        cudaEventRecord(ctx->ev[0], st);
        map_reads_kernel<<<(unsigned)blocks2, (unsigned)(bw2 * 32), 0, st>>>(mb2);
        cudaEventRecord(ctx->ev[1], st);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));
        lra_b200_kernel_stat s3; memset(&s3, 0, sizeof s3); snprintf(s3.name, sizeof s3.name, "map_reads(retry, 8x scratch)"); cudaEventElapsedTime(&s3.ms, ctx->ev[0], ctx->ev[1]);
        s3.jobs = (uint64_t)redo.size(); all.push_back(s3);
      }
    }
  }
  CU(cudaMemcpyAsync(hcur, B[8].p, 32, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (mb.C.prof) {
    std::vector<unsigned long long> hp((size_t)(warps + 4) * lra::mp::kProfStages);
    CU(cudaMemcpy(hp.data(), B[27].p, hp.size() * 8, cudaMemcpyDeviceToHost));
    static const char *nm[lra::mp::kProfStages] = {"minimizers+sort", "CompareLists(global)", "strand+CleanMatches", "LinearExtend#1", "SparseDP#1", "SPLITChain", "Refine_splitchain",
                                                   "Refine_Btwnsplitchain", "LinearExtend#2+Trim", "SparseDP#2+filters", "LocalRefineAlignment(all)", "  AffineOneGapAlign", "  RefineSpace",
                                                   "  SparseDP#3", "output", "phase barriers", "  [all SDP] points+sorts", "  [all SDP] divide", "  [all SDP] ProcessPoint", "    divide: partition", "    divide: unique", "    divide: SS lists", "    divide: sub-problem set-up", "    divide: node bookkeeping"};
    unsigned long long tot[lra::mp::kProfStages] = {0}; unsigned long long all_c = 0;
    for (int wv = 0; wv < warps; wv++) for (int s = 0; s < lra::mp::kProfStages; s++) tot[s] += hp[(size_t)wv * lra::mp::kProfStages + s];
    for (int s = 0; s < 11; s++) all_c += tot[s]; all_c += tot[14] + tot[15];
    fprintf(stderr, "[lra_b200 map profile] %d warps, arena %zu MB/warp, peak %.1f MB; share of worker cycles:\n", warps, per >> 20, (double)hcur[3] / 1e6);
    for (int s = 0; s < 24; s++) fprintf(stderr, "  %-28s %6.2f %%\n", nm[s], all_c ? 100.0 * (double)tot[s] / (double)all_c : 0.0);
  }
  const int S = (int)(hcur[0] >> 40); const unsigned long long NB = hcur[0] & ((1ull << 40) - 1ull);
  const int kerr = (int)(hcur[2] & 0xffffffffull);
  if (kerr & 3) return fail(ctx, LRA_B200_EOVERFLOW, "map_batch: segment / block capacity exceeded (%d segments, %llu blocks)", S, NB);
  // ---- a19 IndelRefineAlignment and a21 CalculateStatistics over all segments
  if ((rc = ensure(ctx, B[11], (size_t)(S + 1) * 8)) || (rc = ensure(ctx, B[12], (size_t)(S + 1) * 4)) || (rc = ensure(ctx, B[13], (size_t)(S + 1) * 4)) ||
      (rc = ensure(ctx, B[14], (size_t)(S + 1) * 4)) || (rc = ensure(ctx, B[15], (size_t)(S + 1) * 4)) || (rc = ensure(ctx, B[16], (size_t)(S + 1) * 4)))
    return rc;
  // MapRead_highacc (Map_highacc.h:716-731): IndelRefineAlignment(endAlign = true), CalculateStatistics, RefineBreakpoint between consecutive segments of an
  // alignment, CalculateStatistics again.  The alignments with several segments are rare (reads across structural variants): their pairs are
  // refined round by round (segment s against s - 1; a segment is the left side of one pair and the right side of the next) and the new block lists are
  // appended behind the IndelRefine output.
  const bool HA = m->opts.HighlyAccurate != 0;
  struct HaAln { int seg0, nseg; };
  std::vector<HaAln> multi;
  std::vector<SegRec> h_seg;
  size_t rbp_slack = 0;
  if (HA && S > 0) {
    std::vector<int> h_nch(n_reads), h_nseg((size_t)n_reads * 4), h_seg0((size_t)n_reads * 4);
    CU(cudaMemcpyAsync(h_nch.data(), B[3].p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_nseg.data(), B[4].p, (size_t)n_reads * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_seg0.data(), B[5].p, (size_t)n_reads * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int r = 0; r < n_reads; r++) for (int a = 0; a < h_nch[r] && a < 4; a++) if (h_nseg[(size_t)r * 4 + a] > 1) multi.push_back(HaAln{h_seg0[(size_t)r * 4 + a], h_nseg[(size_t)r * 4 + a]});
    if (!multi.empty()) {
      h_seg.resize(S);
      CU(cudaMemcpyAsync(h_seg.data(), B[6].p, (size_t)S * sizeof(SegRec), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      for (const HaAln &a : multi) for (int k = 0; k < a.nseg; k++) rbp_slack += 2 * (2 * (size_t)h_seg[a.seg0 + k].blk_cnt + 64 + 2 * (size_t)lra::kRbpCap);
    }
  }
  const size_t ir_cap = (size_t)NB * 2 + (size_t)S * 64 + 1024 + rbp_slack;
  const size_t cig_cap = (size_t)total_bases / 2 + (size_t)NB * 4 + (size_t)S * 16 + 1024;
  if ((rc = ensure(ctx, B[17], (size_t)(S + 1) * 4)) || (rc = ensure(ctx, B[18], (size_t)(S + 1) * 8)) || (rc = ensure(ctx, B[19], ir_cap * 12)) ||
      (rc = ensure(ctx, B[20], (size_t)(S + 1) * 64)) || (rc = ensure(ctx, B[21], (size_t)(S + 1) * 4)) || (rc = ensure(ctx, B[22], (size_t)(S + 2) * 8)) ||
      (rc = ensure(ctx, B[23], cig_cap * 4)) || (rc = ensure(ctx, B[24], (size_t)(S + 1) * sizeof(lra_b200_record))) || (rc = ensure(ctx, B[25], (size_t)n_reads * 16)) ||
      (rc = ensure(ctx, B[26], 64)))
    return rc;
  lra_b200_ir_seg_result ir; memset(&ir, 0, sizeof ir);
  lra_b200_stats_result sr; memset(&sr, 0, sizeof sr);
  if (S > 0) {
    seg_to_ir_kernel<<<(unsigned)((S + 127) / 128), 128, 0, st>>>((const SegRec *)B[6].p, S, (const unsigned long long *)rs->off.p, (const uint32_t *)rs->len.p, Npad, m->d_hdr,
                                                                 (unsigned long long *)B[11].p, (int32_t *)B[12].p, (uint32_t *)B[13].p, (uint32_t *)B[14].p, (int32_t *)B[15].p,
                                                                 (int32_t *)B[16].p);
    ctx->launches++;
    CU(cudaGetLastError());
    lra_b200_ir_segments sg; memset(&sg, 0, sizeof sg);
    sg.blocks_in = (const uint32_t *)B[7].p; sg.blk_off = (const uint64_t *)B[11].p; sg.blk_cnt = (const int32_t *)B[12].p; sg.q_base = (const uint32_t *)B[13].p;
    sg.t_base = (const uint32_t *)B[14].p; sg.read_len = (const int32_t *)B[15].p; sg.contig_len = (const int32_t *)B[16].p; sg.n_blocks_in = NB; sg.n_segments = S;
    sg.refine_band = m->opts.refineBand; sg.match = m->opts.localMatch; sg.mismatch = m->opts.localMismatch; sg.indel = m->opts.localIndel; sg.end_align = HA ? 1 : 0;
    ir.n_blocks = (int32_t *)B[17].p; ir.block_off = (uint64_t *)B[18].p; ir.blocks = (uint32_t *)B[19].p; ir.block_cap = ir_cap;
    if ((rc = lra_b200_indel_refine_batch_device(ctx, rs->reads, m->genome, &sg, &ir))) return rc;
    all.insert(all.end(), ctx->stats.begin(), ctx->stats.end());
    lra_b200_ir_segments s2 = sg;
    s2.blocks_in = ir.blocks; s2.blk_off = ir.block_off; s2.blk_cnt = ir.n_blocks; s2.n_blocks_in = ir.n_blocks_total;
    sr.stats = (int32_t *)B[20].p; sr.value = (float *)B[21].p; sr.cigar_off = (uint64_t *)B[22].p; sr.cigar = (uint32_t *)B[23].p; sr.cigar_cap = cig_cap;
    if ((rc = lra_b200_calc_stats_batch_device(ctx, rs->reads, m->genome, &s2, m->log_lut, &sr))) return rc;
    all.insert(all.end(), ctx->stats.begin(), ctx->stats.end());
    if (HA && !multi.empty()) {
      if ((rc = ensure(ctx, B[29], (size_t)(S + 1) * 64))) return rc;
      CU(cudaMemcpyAsync(B[29].p, B[20].p, (size_t)S * 64, cudaMemcpyDeviceToDevice, st));
      std::vector<int32_t> h_nb(S); std::vector<uint64_t> h_bo(S);
      CU(cudaMemcpyAsync(h_nb.data(), B[17].p, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(h_bo.data(), B[18].p, (size_t)S * 8, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      uint64_t cursor = ir.n_blocks_total;
      lra_b200_seq fwd_v = *rs->reads, rc_v = *rs->reads;
      fwd_v.n = Npad; rc_v.b2 += Npad / 16; rc_v.nm += Npad / 32; rc_v.n = Npad;
      uint32_t *blk = (uint32_t *)B[19].p;
      int max_seg = 0; for (const HaAln &a : multi) max_seg = a.nseg > max_seg ? a.nseg : max_seg;
      for (int j = 1; j < max_seg; j++) {
        std::vector<int> li, ri;
        for (const HaAln &a : multi) if (a.nseg > j && h_nb[a.seg0 + j] > 0 && h_nb[a.seg0 + j - 1] > 0) { li.push_back(a.seg0 + j); ri.push_back(a.seg0 + j - 1); }
        const size_t P = li.size();
        if (P == 0) continue;
        std::vector<uint32_t> lf(P * 3), ll(P * 3), rf(P * 3), rl(P * 3), rlen(P), lcl(P), rcl(P);
        std::vector<uint8_t> ls(P), rsd(P);
        std::vector<uint64_t> roff(P), lco(P), rco(P);
        for (size_t p = 0; p < P; p++) {
          const int l = li[p], r_ = ri[p];
          CU(cudaMemcpyAsync(&lf[p * 3], blk + 3 * h_bo[l], 12, cudaMemcpyDeviceToHost, st));
          CU(cudaMemcpyAsync(&ll[p * 3], blk + 3 * (h_bo[l] + (uint64_t)h_nb[l] - 1), 12, cudaMemcpyDeviceToHost, st));
          CU(cudaMemcpyAsync(&rf[p * 3], blk + 3 * h_bo[r_], 12, cudaMemcpyDeviceToHost, st));
          CU(cudaMemcpyAsync(&rl[p * 3], blk + 3 * (h_bo[r_] + (uint64_t)h_nb[r_] - 1), 12, cudaMemcpyDeviceToHost, st));
          ls[p] = (uint8_t)h_seg[l].strand; rsd[p] = (uint8_t)h_seg[r_].strand;
          roff[p] = rs->h_off[h_seg[l].read]; rlen[p] = rs->h_len[h_seg[l].read];
          lco[p] = m->h_hdr[h_seg[l].chrom]; lcl[p] = (uint32_t)(m->h_hdr[h_seg[l].chrom + 1] - m->h_hdr[h_seg[l].chrom]);
          rco[p] = m->h_hdr[h_seg[r_].chrom]; rcl[p] = (uint32_t)(m->h_hdr[h_seg[r_].chrom + 1] - m->h_hdr[h_seg[r_].chrom]);
        }
        CU(cudaStreamSynchronize(st));
        lra_b200_breakpoints bp; memset(&bp, 0, sizeof bp);
        bp.n_pairs = (int32_t)P; bp.lf = lf.data(); bp.ll = ll.data(); bp.rf = rf.data(); bp.rl = rl.data(); bp.lstrand = ls.data(); bp.rstrand = rsd.data();
        bp.read_off = roff.data(); bp.read_len = rlen.data(); bp.lchrom_off = lco.data(); bp.rchrom_off = rco.data(); bp.lchrom_len = lcl.data(); bp.rchrom_len = rcl.data();
        std::vector<int32_t> mode(P * 2), n_out(P * 2), refined(P);
        std::vector<uint32_t> bound(P * 6), outb(P * 2 * (size_t)lra::kRbpCap * 3);
        lra_b200_breakpoint_result br; br.mode = mode.data(); br.n_out = n_out.data(); br.bound = bound.data(); br.out = outb.data(); br.refined = refined.data();
        if ((rc = lra_b200_refine_breakpoint_batch(ctx, &fwd_v, &rc_v, m->genome, &bp, &br))) return rc;
        all.insert(all.end(), ctx->stats.begin(), ctx->stats.end());
        for (size_t p = 0; p < P; p++) for (int side = 0; side < 2; side++) {
          const int md = mode[2 * p + side], no = n_out[2 * p + side], x = side == 0 ? li[p] : ri[p];
          if (md != 1 && md != 2) continue;
          const uint64_t n_old = (uint64_t)h_nb[x];
          if (cursor + n_old + (uint64_t)no > ir_cap) return fail(ctx, LRA_B200_EOVERFLOW, "map_batch: block capacity exceeded while splicing refined breakpoints");
          uint32_t *dst = blk + 3 * cursor;
          const uint32_t *src = blk + 3 * h_bo[x];
          const uint32_t *nb_ = &outb[(2 * p + side) * (size_t)lra::kRbpCap * 3], *bd = &bound[(2 * p + side) * 3];
          if (md == 1) {
            CU(cudaMemcpyAsync(dst, src, n_old * 12, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(dst + 3 * (n_old - 1), bd, 12, cudaMemcpyHostToDevice, st));
            if (no) CU(cudaMemcpyAsync(dst + 3 * n_old, nb_, (size_t)no * 12, cudaMemcpyHostToDevice, st));
          } else {
            if (no) CU(cudaMemcpyAsync(dst, nb_, (size_t)no * 12, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(dst + 3 * (size_t)no, src, n_old * 12, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(dst + 3 * (size_t)no, bd, 12, cudaMemcpyHostToDevice, st));
          }
          h_bo[x] = cursor; h_nb[x] = (int32_t)(n_old + (uint64_t)no); cursor += n_old + (uint64_t)no;
        }
        CU(cudaStreamSynchronize(st));
      }
      CU(cudaMemcpyAsync(B[17].p, h_nb.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(B[18].p, h_bo.data(), (size_t)S * 8, cudaMemcpyHostToDevice, st));
      CU(cudaStreamSynchronize(st));
      s2.n_blocks_in = cursor;
      if ((rc = lra_b200_calc_stats_batch_device(ctx, rs->reads, m->genome, &s2, m->log_lut, &sr))) return rc;
      all.insert(all.end(), ctx->stats.begin(), ctx->stats.end());
    }
  }
  // SimpleMapQV without bypassClustering reads logf(value / globalK) per segment (Mapping_ultility.h:524, 564): evaluated with the host libm, as in the reference
  if (!m->opts.bypassClustering) {
    if ((rc = ensure(ctx, B[30], (size_t)(S + 1) * 4))) return rc;
    if (S > 0) {
      std::vector<float> hv(S);
      CU(cudaMemcpyAsync(hv.data(), B[21].p, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      for (int i = 0; i < S; i++) hv[i] = hv[i] > 3 ? logf(hv[i] / m->opts.globalK) : 0.0f;
      CU(cudaMemcpyAsync(B[30].p, hv.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
      CU(cudaStreamSynchronize(st));
    }
  }
  // ---- finalize
  CU(cudaMemcpyAsync(B[26].p, m->logf_len, 32, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync((char *)B[26].p + 32, 0, 8, st));
  FinalBatch fb;
  fb.n_reads = n_reads; fb.o = m->opts; fb.read_off = (const unsigned long long *)rs->off.p; fb.read_len = (const uint32_t *)rs->len.p; fb.status = (const int *)B[2].p;
  fb.n_chains = (const int *)B[3].p; fb.chain_nseg = (const int *)B[4].p; fb.chain_seg0 = (const int *)B[5].p; fb.seg = (const SegRec *)B[6].p;
  fb.ir_nblk = (const int32_t *)B[17].p; fb.ir_off = (const unsigned long long *)B[18].p; fb.ir_blocks = (const uint32_t *)B[19].p;
  fb.stats = (const int32_t *)B[20].p; fb.value = (const float *)B[21].p; fb.cigar_off = (const unsigned long long *)B[22].p; fb.logf_len = (const float *)B[26].p;
  fb.stats_first = HA ? (const int32_t *)((!multi.empty() && S > 0) ? B[29].p : B[20].p) : nullptr;
  fb.seg_l = m->opts.bypassClustering ? nullptr : (const float *)B[30].p;
  fb.rec = (lra_b200_record *)B[24].p; fb.rank = (int *)B[25].p; fb.aligned_bases = (unsigned long long *)((char *)B[26].p + 32);
  map_finalize_kernel<<<(unsigned)((n_reads + 127) / 128), 128, 0, st>>>(fb);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
  m->last_S = S; m->last_ncig = sr.n_cigar_total; m->last_kerr = kerr;
  // algorithmic bytes of the whole path per batch: ASCII in, packed forward + reverse strands, one 32-byte sector per global-index probe
  // (2 L / (w + 1) read minimizers), the reference bases and LocalIndex slices under the reads (~ the read bases), CIGAR words and records out
  for (auto &k : all)
    if (!strcmp(k.name, "map_reads"))
      k.algo_bytes = total_bases + 2 * ((total_bases + 3) / 4) + 32ull * (2ull * total_bases / (unsigned long long)(m->opts.globalW + 1)) + (total_bases + 3) / 4 +
                     (unsigned long long)(1.44 * (double)total_bases) + 4ull * sr.n_cigar_total + 64ull * (unsigned long long)S;
  ctx->stats = all;
  if (kerr & 4) return fail(ctx, LRA_B200_EINTERNAL, "map_batch: worker scratch exhausted inside TrimOverlappedAnchors");
  return LRA_B200_OK;
}

// records of the last lra_b200_map_resident call -> caller-owned host buffers
extern "C" int lra_b200_map_download(lra_b200_ctx *ctx, lra_b200_mapper *m, lra_b200_map_result *res) {
  if (!ctx || !m || !res) return fail(ctx, LRA_B200_EINVAL, "map_download: NULL argument");
  if (!res->status || !res->n_aln || !res->aln_nseg || !res->aln_seg0 || !res->aln_rank || !res->records || !res->cigar) return fail(ctx, LRA_B200_EINVAL, "map_download: NULL result array");
  CU(cudaSetDevice(ctx->device));
  res->n_records = 0; res->n_cigar = 0; res->aligned_bases = 0;
  const int n_reads = m->last_reads, S = m->last_S;
  if (n_reads == 0) return LRA_B200_OK;
  cudaStream_t st = ctx->stream;
  DevBuf *B = m->b;
  struct { uint64_t n_cigar_total; } sr = {m->last_ncig};
  if ((uint64_t)S > res->record_cap) { res->n_records = (uint64_t)S; return fail(ctx, LRA_B200_EOVERFLOW, "map_batch: record capacity %llu too small, %d needed", (unsigned long long)res->record_cap, S); }
  CU(cudaMemcpyAsync(res->status, B[2].p, (size_t)n_reads * 4, cudaMemcpyDefault, st));
  CU(cudaMemcpyAsync(res->n_aln, B[3].p, (size_t)n_reads * 4, cudaMemcpyDefault, st));
  CU(cudaMemcpyAsync(res->aln_nseg, B[4].p, (size_t)n_reads * 16, cudaMemcpyDefault, st));
  CU(cudaMemcpyAsync(res->aln_seg0, B[5].p, (size_t)n_reads * 16, cudaMemcpyDefault, st));
  CU(cudaMemcpyAsync(res->aln_rank, B[25].p, (size_t)n_reads * 16, cudaMemcpyDefault, st));
  if (S > 0) CU(cudaMemcpyAsync(res->records, B[24].p, (size_t)S * sizeof(lra_b200_record), cudaMemcpyDefault, st));
  if (sr.n_cigar_total > res->cigar_cap) { res->n_cigar = sr.n_cigar_total; return fail(ctx, LRA_B200_EOVERFLOW, "map_batch: cigar capacity %llu too small, %llu needed", (unsigned long long)res->cigar_cap, (unsigned long long)sr.n_cigar_total); }
  if (sr.n_cigar_total) CU(cudaMemcpyAsync(res->cigar, B[23].p, (size_t)sr.n_cigar_total * 4, cudaMemcpyDefault, st));
  unsigned long long ab = 0;
  CU(cudaMemcpyAsync(&ab, (char *)B[26].p + 32, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  res->n_records = (uint64_t)S; res->n_cigar = sr.n_cigar_total; res->aligned_bases = ab;
  return LRA_B200_OK;
}

extern "C" int lra_b200_map_batch(lra_b200_ctx *ctx, lra_b200_mapper *m, const char *reads_ascii, uint64_t reads_len, const uint64_t *read_off, const uint32_t *read_len,
                                  int32_t n_reads, lra_b200_map_result *res) {
  if (!ctx || !m || !res || n_reads < 0 || (n_reads && (!reads_ascii || !read_off || !read_len))) return fail(ctx, LRA_B200_EINVAL, "map_batch: NULL argument");
  if (!res->status || !res->n_aln || !res->aln_nseg || !res->aln_seg0 || !res->aln_rank || !res->records || !res->cigar) return fail(ctx, LRA_B200_EINVAL, "map_batch: NULL result array");
  if (!m->own) m->own = new lra_b200_readset();
  int rc;
  if ((rc = readset_fill(ctx, m->own, reads_ascii, reads_len, read_off, read_len, n_reads))) return rc;
  if ((rc = lra_b200_map_resident(ctx, m, m->own))) return rc;
  return lra_b200_map_download(ctx, m, res);
}
