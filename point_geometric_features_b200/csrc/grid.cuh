// grid.cuh -- device-side helpers of the uniform grid index.
#pragma once
#include "common.cuh"

namespace pgeof {

// Cell coordinate along one axis.  The SAME monotone function is used when the index
// is built and when a query ball is mapped to a cell range, so coverage arguments only
// need monotonicity, never the nominal boundaries.
__device__ __forceinline__ int cell_coord(float x, float lo, float inv_h, int n)
{
    const float t = __fmul_rn(__fsub_rn(x, lo), inv_h);
    const int c = __float2int_rd(t);   // saturating; NaN -> 0
    return min(max(c, 0), n - 1);
}

__device__ __forceinline__ uint32_t cell_index(const GridView& g, float x, float y, float z)
{
    const int cx = cell_coord(x, g.lo[0], g.inv_hx, g.n[0]);
    const int cy = cell_coord(y, g.lo[1], g.inv_h, g.n[1]);
    const int cz = cell_coord(z, g.lo[2], g.inv_h, g.n[2]);
    return ((uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy) * (uint32_t)g.n[0] + (uint32_t)cx;
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace pgeof
