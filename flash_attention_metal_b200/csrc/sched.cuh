// CTA -> (batch, head, block) mapping shared by the tensor-core kernels.
//
// The hardware hands out CTAs in linear block-id order, so the order of the ids is the schedule.
// Two things pull in opposite directions:
//  * balance: when the work per block differs (causal: block i sees i+1 key tiles) the launch should
//    be longest-processing-time-first over ALL heads, i.e. heads fastest, blocks slowest -- with the
//    per-head order (blocks fastest) the heavy blocks of the last head start when the launch is
//    almost over (measured: 1268 -> 1327 TFLOP/s causal forward at the flagship shape);
//  * L2 locality: the CTAs in flight at one time should stream the K/V (or Q/dO) of few heads --
//    heads fastest over all 16 heads of the flagship shape has 128 MB of K/V in flight and costs the
//    non-causal kernels 4 %.
// So heads are dispatched in groups of `group` heads (sized by the host so that a group's streamed
// tensors fit in a fraction of the L2): within a group heads fastest, blocks slowest; equal-work
// launches use group = 1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fa_internal.h"

namespace fa {

struct BlockCoord {
  int b, h, blk;  // batch, head, block index in dispatch order (0 = first dispatched); b < 0: no work
};

// group <= 1: grid = (n_blocks, H, B), per-head order.
// group  > 1: grid = (group, n_blocks, n_groups): x-fastest dispatch = heads of a group fastest, then
//             blocks, then groups.  (A flat 1-D grid with the same order measured 4 % slower on
//             equal-work launches at D = 64, so the 3-D form is kept everywhere.)
__device__ __forceinline__ BlockCoord decode_block(int group, int n_heads, int H) {
  BlockCoord c;
  if (group <= 1) {
    c.b = blockIdx.z;
    c.h = blockIdx.y;
    c.blk = blockIdx.x;
    return c;
  }
  const int hh = blockIdx.z * group + blockIdx.x;
  c.blk = blockIdx.y;
  c.b = hh < n_heads ? hh / H : -1;  // padding CTAs of a short last group
  c.h = hh - c.b * H;
  return c;
}

// heads per dispatch group: all work equal -> 1; otherwise as many heads as stream <= 48 MB
inline int dispatch_group(bool uneven_work, int64_t streamed_bytes_per_head, int n_heads) {
  if (!uneven_work) return 1;
  const int64_t budget_mb = l2_group_mb();  // 48 by default (fa_internal.h)
  int64_t g = (budget_mb << 20) / (streamed_bytes_per_head > 0 ? streamed_bytes_per_head : 1);
  if (g < 1) g = 1;
  if (g > n_heads) g = n_heads;
  const int64_t n_groups = (n_heads + g - 1) / g;  // equal groups: a short last group would start
  return (int)((n_heads + n_groups - 1) / n_groups);  // its heavy blocks when the launch is nearly over
}

// launch geometry for decode_block
inline dim3 dispatch_grid(int group, int n_blocks, int H, int B) {
  if (group <= 1) return dim3((unsigned)n_blocks, (unsigned)H, (unsigned)B);
  return dim3((unsigned)group, (unsigned)n_blocks, (unsigned)((H * B + group - 1) / group));
}

}  // namespace fa
