// g2_walk_d4.cu — the walk kernels of N_GRAVS = 4 (one translation unit per N_GRAVS: they compile in parallel).
#include "g2_walk_kernel.cuh"

int g2_launch_walk_d4(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int stats)
{
  return dispatch_walk<4>(c, A, grid, smem, sr, periodic, unequal, stock, accd, stats);
}
