// trrt_device.cuh -- device-side building blocks shared by the kernels in
// thetarrt.cu: lane groups, bit-packed occupancy grid, line of sight.
//
// Data layout in HBM
//   grid   : uint32 words, bit (x&31) of word [y*wpr + (x>>5)], 1 = free
//            (replaces builtins.imarray, main.py:42; search.py:17-33)
//   trees  : SoA float64 x[K], y[K], theta[K] per query (rrt.py G keys in
//            insertion order), int32 parent[K]
//   samples: int32 (x,y) pairs + float64 theta per iteration (rrt.py:144)
// All arithmetic is fp64 with contraction disabled (-fmad=false); fma() is
// used only where the reference's compiled dependencies contract (see
// oracle/trrt_oracle.c header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "trrt_libm.h"

// Bounds assertions of the checked build (-DTRRT_CHECKED, profiles/tools/checked_build.sh): every index the kernels derive
// from data before a store -- node rows, log positions, partial-minimum slots, heap positions, pixel-list positions -- is
// tested and a violation traps (the launch fails with an error instead of corrupting memory).  compute-sanitizer is closed
// on this GPU pool, so the GPU test-suite is run once per round against this build instead (profiles/r2/NOTES.md).
#ifdef TRRT_CHECKED
#include <stdio.h>
#define TRRT_CHECK(c) do { if (!(c)) { printf("TRRT_CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define TRRT_CHECK(c) do { } while (0)
#endif

namespace trrt {

// ---------------------------------------------------------------------------
// Lane groups: G consecutive lanes of a warp cooperate on one query.
// ---------------------------------------------------------------------------
template <int G>
struct Group {
    static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group size");
    unsigned mask; // lanes of this group
    int gl;        // lane index inside the group
    __device__ __forceinline__ Group() {
        int lane = threadIdx.x & 31;
        gl = lane & (G - 1);
        mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    }
    __device__ __forceinline__ void sync() const { if (G > 1) __syncwarp(mask); }
    __device__ __forceinline__ bool any(bool p) const {
        if (G == 1) return p;
        return (__ballot_sync(mask, p) & mask) != 0u;
    }
    __device__ __forceinline__ unsigned ballot(bool p) const {
        if (G == 1) return p ? 1u : 0u;
        return (__ballot_sync(mask, p) & mask) >> ((threadIdx.x & 31) & ~(G - 1));
    }
    template <typename T>
    __device__ __forceinline__ T bcast(T v, int src_gl) const {
        if (G == 1) return v;
        return __shfl_sync(mask, v, src_gl, G);
    }
    // lexicographic (d, i) minimum across the group; every lane gets the result
    __device__ __forceinline__ void min_di(double &d, int &i) const {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            double od = __shfl_xor_sync(mask, d, off, G);
            int oi = __shfl_xor_sync(mask, i, off, G);
            if (od < d || (od == d && oi < i)) { d = od; i = oi; }
        }
    }
    __device__ __forceinline__ unsigned long long sum(unsigned long long v) const {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, G);
        return v;
    }
};

// ---------------------------------------------------------------------------
// Occupancy grid (search.py:17-33).  Square maps only (see thetarrt.h).
// ---------------------------------------------------------------------------
struct Grid {
    const uint32_t *bits;
    int W, H, wpr;
    __device__ __forceinline__ bool inb(long long x, long long y) const {
        // search.valid: 0 <= int(x) < shape[0], 0 <= int(y) < shape[1]; H == W
        return x >= 0 && y >= 0 && x < (long long)H && y < (long long)W;
    }
    // in-bounds pixel -> free?
    __device__ __forceinline__ bool free_nb(int x, int y) const {
        return (__ldg(bits + (size_t)y * wpr + (x >> 5)) >> (x & 31)) & 1u;
    }
    // search.freespace: out of bounds counts as blocked
    __device__ __forceinline__ bool freespace(long long x, long long y) const {
        return inb(x, y) && free_nb((int)x, (int)y);
    }
};

__device__ __forceinline__ long long trunc_ll(double v) { return (long long)v; } // Python int(): toward zero

// ---------------------------------------------------------------------------
// search.lineofsight (search.py:35-94), evaluated by a lane group.
// Bresenham's running error has the closed form
//     minor(i) = minor0 + step * floor((2*dmin*i + dmaj - 1) / (2*dmaj))
// for the i-th pixel along the driving axis (derivation in DESIGN.md), so the
// pixels are independent and lanes test them in parallel.  Endpoints are
// pixels of the line, and the line stays inside their bounding box, so one
// out-of-bounds endpoint decides the result without walking.
// pixels_tested (optional) accumulates max(|dx|,|dy|)+1 for in-bounds rays.
// ---------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ bool los_group(const Group<G> &g, const Grid &m, long long ax, long long ay, long long bx,
                                          long long by, unsigned long long *pixels_tested = nullptr) {
    if (!m.inb(ax, ay) || !m.inb(bx, by)) return false;
    int x0 = (int)ax, y0 = (int)ay, x1 = (int)bx, y1 = (int)by;
    int adx = abs(x1 - x0), ady = abs(y1 - y0);
    bool low = ady < adx; // search.py:47
    if (low ? (x0 > x1) : (y0 > y1)) { int t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
    unsigned dmaj = low ? (unsigned)adx : (unsigned)ady;
    unsigned dmin = low ? (unsigned)ady : (unsigned)adx;
    int step = low ? ((y1 < y0) ? -1 : 1) : ((x1 < x0) ? -1 : 1);
    unsigned n = dmaj + 1u;
    if (pixels_tested) *pixels_tested += n;
    // lane l tests pixels l, l+G, ...: the minor offset s_i = floor((2*dmin*i + dmaj - 1) / (2*dmaj)) is advanced by
    // G pixels per step with a quotient / remainder pair (one division per ray, none per pixel)
    const unsigned den = dmaj ? 2u * dmaj : 1u;
    const unsigned num0 = 2u * dmin * (unsigned)g.gl + dmaj - (dmaj ? 1u : 0u);
    unsigned s = num0 / den, rem = num0 - s * den;
    const unsigned inc = 2u * dmin * (unsigned)G, q = inc / den, r = inc - q * den;
    for (unsigned base = 0; base < n; base += G) {
        unsigned i = base + (unsigned)g.gl;
        bool blocked = false;
        if (i < n) {
            int px = low ? x0 + (int)i : x0 + step * (int)s;
            int py = low ? y0 + step * (int)s : y0 + (int)i;
            blocked = !m.free_nb(px, py);
        }
        if (g.any(blocked)) return false;
        s += q; rem += r;
        if (rem >= den) { rem -= den; s++; }
    }
    return true;
}

} // namespace trrt
