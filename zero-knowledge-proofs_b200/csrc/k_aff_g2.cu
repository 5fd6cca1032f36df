// hot kernel, affine variant: bucket accumulation with block-shared inversions (g2)
#include "kernel_impl.cuh"
namespace g16 {
template size_t k_affine_scratch_words<Fq2>(size_t, size_t, uint32_t);
template void k_accumulate_affine<Fq2>(stream_t, size_t, const uint32_t *, const uint32_t *, const WorkItem *, const uint32_t *,
                                      const uint32_t *, uint32_t, uint32_t *, size_t, size_t, uint32_t *);
}
