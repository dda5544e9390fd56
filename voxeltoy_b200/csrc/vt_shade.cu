// vt_shade.cu -- the translation unit of wf_shade_kernel (csrc/vt_wavefront.cuh), compiled with VT_COMPACT_MATH: IEEE division
// and pow are called, not inlined (see vt_math.cuh). Same arithmetic, same results; 40 % less code in the one kernel that is
// bound by instruction fetch. vt_api.cu reaches the kernel through the two functions below.
#define VT_COMPACT_MATH 1
#include "vt_wavefront.cuh"

namespace vt {

cudaError_t wf_shade_blocks_per_sm(bool count, int* blocks)
{
    return count ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, wf_shade_kernel<true>, VT_WF_SHADE_THREADS, 0)
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, wf_shade_kernel<false>, VT_WF_SHADE_THREADS, 0);
}

void wf_shade_launch(bool count, unsigned int blocks, cudaStream_t st, const Volume& V, const Frame& F, const WfState& S, int gen,
                     WfCounts* cnt, Counters* counters)
{
    if (count) wf_shade_kernel<true><<<blocks, VT_WF_SHADE_THREADS, 0, st>>>(V, F, S, gen, cnt, counters);
    else wf_shade_kernel<false><<<blocks, VT_WF_SHADE_THREADS, 0, st>>>(V, F, S, gen, cnt, counters);
}

} // namespace vt
