// The VCF genotype parser as a reusable object (vcf_parse.cu): one parse call's parameters, the
// per-record work and the final gather -- used by sai_vcf_parse_gt (text in memory) and by
// sai_bgzf_parse_gt (bgzf.cu: groups of bgzip blocks inflated and parsed by the same thread).
#pragma once
#include <stdint.h>

#include <vector>

namespace sai {

struct KeptLine {
  const char* samples;  // first character after the FORMAT column
  const char* line;     // start of the record
  int32_t gt_index;     // position of GT among the ':'-separated FORMAT keys
  int32_t pos;
  bool flip;
};

struct GtParser {
  struct Run {
    int col0, out0, len;
  };
  struct Scratch {  // per thread: alleles / diploid sums of every field of a regular record
    std::vector<int8_t> a0, a1, sum2;
  };
  struct SegOut {  // what one segment (or block group) keeps, in file order
    std::vector<KeptLine> kept;
    std::vector<int8_t> rows;  // kept.size() rows of n_out values
  };
  // false: bad sample column / ploidy
  bool init(const char* chrom, int64_t start, int64_t end, const int32_t* sample_column, const int32_t* sample_ploidy,
            int32_t n_out, const int32_t* anc_pos, const char* anc_allele, int64_t n_anc);
  void parse_record(const KeptLine& K, const char* lend, int8_t* row, Scratch& sc) const;
  void scan(const char* p, const char* send, SegOut& so, Scratch& sc) const;
  int64_t gather(const std::vector<SegOut>& seg_out, int32_t* out_pos, int8_t* out_gt, int64_t row_stride, int64_t rows_cap,
                 int n_threads, const KeptLine** first_dropped) const;

  const char* chrom = nullptr;
  size_t chrom_len = 0;
  int64_t start = 1, end = 0;
  bool region = false;
  const int32_t* sample_column = nullptr;
  const int32_t* sample_ploidy = nullptr;
  int32_t n_out = 0;
  const int32_t* anc_pos = nullptr;
  const char* anc_allele = nullptr;
  int64_t n_anc = 0;
  std::vector<int> order;
  std::vector<Run> runs;
  bool all_diploid = true, by_runs = false;
};

}  // namespace sai
