// Host-side VCF genotype parser (N2: ingest).  No device code in this file.
//
// Replaces, for the scoring path, the per-population `allel.read_vcf(...,
// numbers={"GT": ploidy}, region=...)` calls of read_geno_data
// (sai/utils/utils.py:78-186), the ancestral-allele filter / flip of
// check_anc_allele + flip_snps (utils.py:492-555) and the allele sum of
// reshape_genotypes(is_phased=False) (utils.py:405-410) by ONE pass over the
// VCF text that writes, per kept record, the position and one int8 allele sum
// per requested (sample column, ploidy) pair:
//   * GT is cut or padded with missing (-1) to `ploidy` alleles, "." = -1;
//   * without an ancestral-allele table every record of the region is kept;
//     with one, records whose position is not in the table or whose ancestral
//     allele is neither REF nor the first ALT are dropped, and where it equals
//     ALT every allele a becomes |a - 1| (so a missing allele becomes 2);
//   * the value written is the sum of the (possibly flipped) alleles.
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"
#include "vcf_parse.h"
#include "vcf_simd.h"

namespace sai {


static inline const char* next_tab(const char* p, const char* end) {
  const void* t = memchr(p, '\t', (size_t)(end - p));
  return t ? static_cast<const char*>(t) : end;
}

// Parses the GT of the sample field starting at f (field ends at a tab or at
// lend) into the allele sum (see file header); returns the end of the field.
static inline const char* gt_sum(const char* f, const char* lend, int gt_index, int ploidy,
                                 bool flip, int8_t* out) {
  const char* p = f;
  // skip to the GT sub-field
  for (int k = 0; k < gt_index; ++k) {
    while (p < lend && *p != ':' && *p != '\t') ++p;
    if (p < lend && *p == ':') ++p;
  }
  int sum = 0, n = 0;
  bool open = p < lend && *p != '\t' && *p != ':';  // still inside the GT sub-field
  while (n < ploidy) {
    int a = -1;
    if (open) {
      int v = 0;
      bool digits = false, junk = false;
      while (p < lend) {
        const char c = *p;
        if (c == '/' || c == '|' || c == ':' || c == '\t') break;
        if (c >= '0' && c <= '9' && !junk) {
          v = v * 10 + (c - '0');
          digits = true;
        } else {
          junk = true;  // "." or anything else: missing
        }
        ++p;
      }
      if (digits && !junk) a = v;
      if (p < lend && (*p == '/' || *p == '|'))
        ++p;  // next allele token (possibly empty)
      else
        open = false;
    }
    if (flip) a = a > 0 ? a - 1 : 1 - a;  // |a - 1|
    sum += a;
    ++n;
  }
  *out = (int8_t)(sum < -128 ? -128 : (sum > 127 ? 127 : sum));
  while (p < lend && *p != '\t') ++p;
  return p;
}

// allele of a one-character token: digit, "." (missing = -1) or kBadAllele
constexpr int kBadAllele = -100;
static inline int allele_of(char c) {
  if (c >= '0' && c <= '9') return c - '0';
  return c == '.' ? -1 : kBadAllele;
}

// ---------------------------------------------------------------------------------------------
// GtParser (vcf_parse.h): everything one parse call needs to know, the per-record work and the
// final gather -- shared by the text entry point below and the fused bgzip one (bgzf.cu).

bool GtParser::init(const char* chrom_, int64_t start_, int64_t end_, const int32_t* sample_column_,
                    const int32_t* sample_ploidy_, int32_t n_out_, const int32_t* anc_pos_, const char* anc_allele_,
                    int64_t n_anc_) {
  chrom = chrom_;
  chrom_len = strlen(chrom_);
  start = start_;
  end = end_;
  region = start_ <= end_;
  sample_column = sample_column_;
  sample_ploidy = sample_ploidy_;
  n_out = n_out_;
  anc_pos = anc_pos_;
  anc_allele = anc_allele_;
  n_anc = n_anc_;
  // columns sorted so that every line is walked once, left to right
  order.resize(n_out);
  for (int i = 0; i < n_out; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return sample_column[a] < sample_column[b]; });
  for (int i = 0; i < n_out; ++i)
    if (sample_column[i] < 0 || sample_ploidy[i] < 1 || sample_ploidy[i] > 16) return false;
  // Diploid requests over runs of consecutive sample columns (a population's individuals usually sit
  // next to each other in the file): a regular record then needs the allele sums of all its fields
  // (one vector sweep, vcf_simd.cpp) and one block copy per run.
  runs.clear();
  all_diploid = true;
  for (int o = 0; o < n_out; ++o) {
    if (sample_ploidy[o] != 2) all_diploid = false;
    if (!runs.empty() && sample_column[o] == runs.back().col0 + runs.back().len && o == runs.back().out0 + runs.back().len)
      ++runs.back().len;
    else
      runs.push_back(Run{sample_column[o], o, 1});
  }
  by_runs = all_diploid && (int64_t)runs.size() * 8 <= n_out;
  return true;
}

// One record -> one output row.
void GtParser::parse_record(const KeptLine& K, const char* lend, int8_t* row, Scratch& sc) const {
  // Regular diploid record (every field `x|y` / `x/y`): all fields are converted 16 at a time,
  // then the requested columns are picked with their own ploidy -- cut / padded with missing
  // alleles and flipped exactly like gt_sum does field by field.
  if (K.gt_index == 0) {
    const int64_t max_fields = (lend - K.samples + 1) / 4 + 1;
    if ((int64_t)sc.sum2.size() < max_fields) {
      sc.a0.resize(max_fields);
      sc.a1.resize(max_fields);
      sc.sum2.resize(max_fields);
    }
    int64_t nf = 0;
    if (all_diploid) {
      if (vcf_regular_diploid_sum(K.samples, lend, K.flip, sc.sum2.data(), (int64_t)sc.sum2.size(), &nf)) {
        const int8_t* s2 = sc.sum2.data();
        if (by_runs) {
          for (const Run& r : runs) {
            const int64_t have = std::max<int64_t>(0, std::min<int64_t>(r.len, nf - r.col0));
            if (have > 0) memcpy(row + r.out0, s2 + r.col0, (size_t)have);
            if (have < r.len) memset(row + r.out0 + have, -2, (size_t)(r.len - have));  // column absent: both alleles missing
          }
        } else {
          for (int o = 0; o < n_out; ++o) {
            const int col = sample_column[o];
            row[o] = col < nf ? s2[col] : (int8_t)-2;
          }
        }
        return;
      }
    } else if (vcf_regular_diploid(K.samples, lend, sc.a0.data(), sc.a1.data(), (int64_t)sc.a0.size(), &nf)) {
      const int8_t *p0 = sc.a0.data(), *p1 = sc.a1.data();
      for (int o = 0; o < n_out; ++o) {
        const int col = sample_column[o], ploidy = sample_ploidy[o];
        if (col >= nf) {
          row[o] = (int8_t)(-ploidy);  // column absent: all alleles missing
          continue;
        }
        int x0 = p0[col], x1 = ploidy >= 2 ? p1[col] : 0;
        int rest = ploidy > 2 ? ploidy - 2 : 0;  // alleles beyond the field: missing (-1)
        if (K.flip) {
          x0 = x0 > 0 ? x0 - 1 : 1 - x0;
          if (ploidy >= 2) x1 = x1 > 0 ? x1 - 1 : 1 - x1;
          rest *= -2;
        }
        const int sum = x0 + x1 - rest;
        row[o] = (int8_t)(sum < -128 ? -128 : (sum > 127 ? 127 : sum));
      }
      return;
    }
  }
  const char* f = K.samples;  // start of sample column `col`
  const char* fe = nullptr;   // end of the field at f (its tab or lend) when already known
  int col = 0;
  bool have = f < lend;  // a field exists at f
  for (int oi = 0; oi < n_out; ++oi) {
    const int o = order[oi];
    const int want = sample_column[o];
    while (col < want && have) {
      if (fe) {
        f = fe;
        fe = nullptr;
      }
      while (f < lend && *f != '\t') ++f;
      if (f < lend) {
        ++f;
        ++col;
      } else {
        have = false;
      }
    }
    if (col != want || !have) {
      row[o] = (int8_t)(-sample_ploidy[o]);  // column absent: all alleles missing
      continue;
    }
    // fast paths for the overwhelmingly common fields "a|b" / "a/b" (diploid) and "a"
    // (haploid) with single-character alleles and GT first in FORMAT
    const int ploidy = sample_ploidy[o];
    if (K.gt_index == 0 && ploidy <= 2) {
      const int len = ploidy == 2 ? 3 : 1;
      if (f + len <= lend) {
        const char e = f + len < lend ? f[len] : '\t';
        const int a0 = allele_of(f[0]);
        const int a1 = ploidy == 2 ? allele_of(f[2]) : 0;
        const bool sep = ploidy == 1 || f[1] == '|' || f[1] == '/';
        if (sep && a0 != kBadAllele && a1 != kBadAllele && (e == '\t' || e == ':')) {
          int x0 = a0, x1 = a1;
          if (K.flip) {
            x0 = x0 > 0 ? x0 - 1 : 1 - x0;
            x1 = x1 > 0 ? x1 - 1 : 1 - x1;
          }
          row[o] = (int8_t)(ploidy == 2 ? x0 + x1 : x0);
          if (e == '\t') fe = f + len;  // the field ends right here
          continue;
        }
      }
    }
    fe = gt_sum(f, lend, K.gt_index, ploidy, K.flip, &row[o]);  // f stays: a column may be requested twice
  }
}

// Complete lines of [p, send): record filter, flip decision and -- while the line is still in this
// core's cache -- its output row, appended to `so`.
void GtParser::scan(const char* p, const char* send, SegOut& so, Scratch& sc) const {
  std::vector<KeptLine>& kept = so.kept;
  std::vector<int8_t>& rows = so.rows;
  while (p < send) {
    const void* nl = memchr(p, '\n', (size_t)(send - p));
    if (!nl) break;
    const char* lend = static_cast<const char*>(nl);
    const char* line = p;
    p = lend + 1;
    if (lend > line && lend[-1] == '\r') --lend;
    if (line == lend || *line == '#') continue;
    const char* t0 = next_tab(line, lend);  // CHROM
    if ((size_t)(t0 - line) != chrom_len || memcmp(line, chrom, chrom_len) != 0) continue;
    const char* f = t0 + 1;
    const char* t1 = next_tab(f, lend);  // POS
    int64_t pos = 0;
    for (const char* q = f; q < t1; ++q) pos = pos * 10 + (*q - '0');
    if (region && (pos < start || pos > end)) continue;
    const char* t2 = next_tab(t1 + 1, lend);  // ID
    const char* ref = t2 + 1;
    const char* t3 = next_tab(ref, lend);  // REF
    const char* alt = t3 + 1;
    const char* t4 = next_tab(alt, lend);  // ALT
    const void* comma = memchr(alt, ',', (size_t)(t4 - alt));
    const char* alt_end = comma ? static_cast<const char*>(comma) : t4;  // alt_number=1: first ALT
    bool flip = false;
    if (n_anc > 0) {
      const int32_t* it = std::lower_bound(anc_pos, anc_pos + n_anc, (int32_t)pos);
      if (it == anc_pos + n_anc || *it != (int32_t)pos) continue;  // no ancestral allele: dropped
      const char* a = anc_allele + 8 * (it - anc_pos);
      const size_t alen = strnlen(a, 8);
      const bool is_ref = (size_t)(t3 - ref) == alen && memcmp(a, ref, alen) == 0;
      const bool is_alt = (size_t)(alt_end - alt) == alen && memcmp(a, alt, alen) == 0;
      if (!is_ref && !is_alt) continue;
      flip = is_alt;  // utils.py:523-524
    }
    const char* t5 = next_tab(t4 + 1, lend);   // QUAL
    const char* t6 = next_tab(t5 + 1, lend);   // FILTER
    const char* t7 = next_tab(t6 + 1, lend);   // INFO
    const char* fmt = t7 + 1;
    const char* t8 = next_tab(fmt, lend);      // FORMAT
    int gi = 0;
    {
      int k = 0;
      const char* q = fmt;
      while (q < t8) {
        const void* c = memchr(q, ':', (size_t)(t8 - q));
        const char* ke = c ? static_cast<const char*>(c) : t8;
        if (ke - q == 2 && q[0] == 'G' && q[1] == 'T') {
          gi = k;
          break;
        }
        q = ke + 1;
        ++k;
      }
    }
    if (t8 >= lend) continue;  // no sample columns
    const KeptLine K{t8 + 1, line, gi, (int32_t)pos, flip};
    if (kept.empty()) {  // first kept line of the segment: room for a segment of lines like this one
      const size_t est = (size_t)((send - line) / std::max<int64_t>(1, p - line) + 2);
      kept.reserve(est);
      rows.reserve(est * (size_t)n_out);
    }
    kept.push_back(K);
    rows.resize(rows.size() + (size_t)n_out);
    parse_record(K, lend, rows.data() + rows.size() - (size_t)n_out, sc);
  }
}

// Rows of the segments, in order, into the caller's arrays (at most rows_cap; *first_dropped = the
// record in front of which a full output stopped, else NULL).  Returns the number of rows.
int64_t GtParser::gather(const std::vector<SegOut>& seg_out, int32_t* out_pos, int8_t* out_gt, int64_t row_stride,
                         int64_t rows_cap, int n_threads, const KeptLine** first_dropped) const {
  const int n_seg = (int)seg_out.size();
  std::vector<int64_t> first(n_seg + 1, 0);
  for (int si = 0; si < n_seg; ++si) first[si + 1] = first[si] + (int64_t)seg_out[si].kept.size();
  const int64_t n_rows = std::min<int64_t>(first[n_seg], rows_cap);
  *first_dropped = nullptr;
  if (first[n_seg] > rows_cap) {
    int si = 0;
    while (first[si + 1] <= rows_cap) ++si;
    *first_dropped = &seg_out[si].kept[rows_cap - first[si]];
  }
  if (n_rows == 0) return 0;
  std::atomic<int> next_seg{0};
  auto work = [&]() {
    for (int si = next_seg.fetch_add(1); si < n_seg; si = next_seg.fetch_add(1)) {
      const SegOut& so = seg_out[si];
      const int64_t take = std::min<int64_t>((int64_t)so.kept.size(), n_rows - first[si]);
      if (take <= 0) continue;
      for (int64_t r = 0; r < take; ++r) out_pos[first[si] + r] = so.kept[r].pos;
      if (row_stride == n_out) {
        memcpy(out_gt + first[si] * row_stride, so.rows.data(), (size_t)take * n_out);
      } else {
        for (int64_t r = 0; r < take; ++r) memcpy(out_gt + (first[si] + r) * row_stride, so.rows.data() + r * n_out, (size_t)n_out);
      }
    }
  };
  const int nt = std::max(1, std::min(n_threads, n_seg));
  std::vector<std::thread> th;
  for (int i = 1; i < nt; ++i) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return n_rows;
}

}  // namespace sai

using namespace sai;

extern "C" int64_t sai_vcf_parse_gt(const char* text, int64_t len, const char* chrom, int64_t start,
                                    int64_t end, const int32_t* sample_column,
                                    const int32_t* sample_ploidy, int32_t n_out,
                                    const int32_t* anc_pos, const char* anc_allele, int64_t n_anc,
                                    int32_t* out_pos, int8_t* out_gt, int64_t row_stride,
                                    int64_t rows_cap, int64_t* bytes_consumed, int32_t n_threads) {
  if (!text || len < 0 || !chrom || !sample_column || !sample_ploidy || n_out < 1 || !out_pos ||
      !out_gt || row_stride < n_out || rows_cap < 0 || !bytes_consumed || (n_anc > 0 && (!anc_pos || !anc_allele))) {
    set_error("sai_vcf_parse_gt: bad argument");
    return SAI_E_ARG;
  }
  GtParser P;
  if (!P.init(chrom, start, end, sample_column, sample_ploidy, n_out, anc_pos, anc_allele, n_anc)) {
    set_error("sai_vcf_parse_gt: bad sample column / ploidy");
    return SAI_E_ARG;
  }
  // ---- one pass, parallel over byte segments cut at line starts; the rows go to per-segment
  //      buffers first (how many rows the earlier segments keep is not known yet) ----
  const void* last_nl = len > 0 ? memrchr(text, '\n', (size_t)len) : nullptr;
  const char* const complete_end = last_nl ? static_cast<const char*>(last_nl) + 1 : text;  // incomplete last line: next call
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  // a few segments per thread: lines are kept unevenly (region filter, other chromosomes)
  const int64_t seg_bytes = 1 << 20;
  const int n_seg = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_threads * 4, (complete_end - text) / seg_bytes));
  std::vector<const char*> seg(n_seg + 1);
  seg[0] = text;
  seg[n_seg] = complete_end;
  for (int i = 1; i < n_seg; ++i) {
    const char* guess = text + (complete_end - text) / n_seg * i;
    if (guess < seg[i - 1]) guess = seg[i - 1];
    const void* nl = guess < complete_end ? memchr(guess, '\n', (size_t)(complete_end - guess)) : nullptr;
    seg[i] = nl ? static_cast<const char*>(nl) + 1 : complete_end;
  }
  std::vector<GtParser::SegOut> seg_out(n_seg);
  {
    std::atomic<int> next_seg{0};
    auto worker = [&]() {
      GtParser::Scratch sc;
      for (int si = next_seg.fetch_add(1); si < n_seg; si = next_seg.fetch_add(1)) P.scan(seg[si], seg[si + 1], seg_out[si], sc);
    };
    const int nt = std::min(n_threads, n_seg);
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
  }
  const KeptLine* dropped = nullptr;
  const int64_t n_rows = P.gather(seg_out, out_pos, out_gt, row_stride, rows_cap, n_threads, &dropped);
  *bytes_consumed = (int64_t)((dropped ? dropped->line : complete_end) - text);
  return n_rows;
}


// First and last POS (in file order) and the number of records of `chrom` among the complete
// lines of `text` -- what ChunkGenerator.__init__ finds by iterating pysam's fetch(chrom)
// (sai/generators/chunk_generator.py:64-76).  Parallel over byte segments cut at line starts.
extern "C" int sai_vcf_chrom_span(const char* text, int64_t len, const char* chrom, int64_t* first,
                                  int64_t* last, int64_t* n_records, int64_t* bytes_consumed,
                                  int32_t n_threads) {
  if (!text || len < 0 || !chrom || !first || !last || !n_records || !bytes_consumed) {
    set_error("sai_vcf_chrom_span: bad argument");
    return SAI_E_ARG;
  }
  const size_t chrom_len = strlen(chrom);
  const void* last_nl = len > 0 ? memrchr(text, '\n', (size_t)len) : nullptr;
  const char* const complete_end = last_nl ? static_cast<const char*>(last_nl) + 1 : text;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const int n_seg = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, (complete_end - text) / (1 << 20)));
  std::vector<const char*> seg(n_seg + 1);
  seg[0] = text;
  seg[n_seg] = complete_end;
  for (int i = 1; i < n_seg; ++i) {
    const char* guess = text + (complete_end - text) / n_seg * i;
    if (guess < seg[i - 1]) guess = seg[i - 1];
    const void* nl = guess < complete_end ? memchr(guess, '\n', (size_t)(complete_end - guess)) : nullptr;
    seg[i] = nl ? static_cast<const char*>(nl) + 1 : complete_end;
  }
  struct Span {
    int64_t first = -1, last = -1, n = 0;
  };
  std::vector<Span> out(n_seg);
  auto scan = [&](int si) {
    Span sp;
    const char* p = seg[si];
    const char* const send = seg[si + 1];
    while (p < send) {
      const void* nl = memchr(p, '\n', (size_t)(send - p));
      if (!nl) break;
      const char* lend = static_cast<const char*>(nl);
      const char* line = p;
      p = lend + 1;
      if (line == lend || *line == '#') continue;
      if ((size_t)(lend - line) <= chrom_len || memcmp(line, chrom, chrom_len) != 0 || line[chrom_len] != '\t') continue;
      int64_t pos = 0;
      for (const char* q = line + chrom_len + 1; q < lend && *q >= '0' && *q <= '9'; ++q) pos = pos * 10 + (*q - '0');
      if (sp.first < 0) sp.first = pos;
      sp.last = pos;
      ++sp.n;
    }
    out[si] = sp;
  };
  if (n_seg == 1) {
    scan(0);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_seg; ++i) th.emplace_back(scan, i);
    for (auto& t : th) t.join();
  }
  *first = *last = -1;
  *n_records = 0;
  for (const Span& sp : out) {
    if (sp.n == 0) continue;
    if (*first < 0) *first = sp.first;
    *last = sp.last;
    *n_records += sp.n;
  }
  *bytes_consumed = (int64_t)(complete_end - text);
  return SAI_OK;
}
