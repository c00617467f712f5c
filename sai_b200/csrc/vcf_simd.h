// Fast path of the host VCF genotype parser for regular diploid records (vcf_simd.cpp).
#pragma once
#include <stdint.h>

namespace sai {
// Sample region [s, lend) of one record.  Regular = every field is `x|y` / `x/y` (x, y one digit or
// "."), single tabs between fields: writes the alleles (digit value, -1 for ".") of field i to
// a0[i], a1[i], sets *n_fields, returns true.  Otherwise returns false (outputs unspecified).
bool vcf_regular_diploid(const char* s, const char* lend, int8_t* a0, int8_t* a1, int64_t cap, int64_t* n_fields);
// Same recognition, but only the diploid allele SUM of every field is written (sum2[i] = a0 + a1,
// or |a0 - 1| + |a1 - 1| when `flip`): what a request for ploidy 2 needs, in one sweep.
bool vcf_regular_diploid_sum(const char* s, const char* lend, bool flip, int8_t* sum2, int64_t cap, int64_t* n_fields);
}  // namespace sai
