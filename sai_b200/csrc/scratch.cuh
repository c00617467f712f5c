// Stream-ordered scratch memory for kernels that need a temporary between two launches
// (per-site products of N3, per-site distances of N4).  A private memory pool per device whose
// release threshold is "never": after the first call an allocation is a pointer bump on the
// stream, and no global allocator state of the host application (the default pool, torch's
// caching allocator) is touched.
#pragma once
#include <cuda_runtime.h>

namespace sai {
int scratch_alloc(void** out, size_t bytes, cudaStream_t st);  // SAI_OK or SAI_E_NOMEM / SAI_E_CUDA
int scratch_free(void* p, cudaStream_t st);
}  // namespace sai
