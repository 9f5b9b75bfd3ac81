// nbr.cuh -- radius-neighbourhood iteration over the voxel table (device side).
//
// A query sphere (p, R) touches a rectangle of (iy,iz) rows; inside one row the needed cells are a
// contiguous range of the cell-sorted float4 array.  A CTA (or warp) first turns the rows into a
// list of non-empty SEGMENTS (start,len) + exclusive prefix in shared memory, then strides over the
// flattened candidate index so that consecutive lanes read consecutive float4s (coalesced, 16 B
// vector loads) regardless of how ragged the segments are.
#pragma once
#include "common.cuh"

namespace bshot {

__host__ __device__ __forceinline__ int cell_coord(float v, float o, float inv) {
    return (int)floorf((v - o) * inv);
}

struct RowRange {
    int iy_lo, iz_lo, ny_span, nrows;
    float inv_span;
};

// row r of the rectangle -> (iy, iz) without an integer division: (r + 0.5) / span is at least
// 0.5 / span away from an integer, far more than the fp32 error for r < 2^16
__device__ __forceinline__ void row_coords(const RowRange& rr, int r, int& iy, int& iz) {
    const int q = (int)(((float)r + 0.5f) * rr.inv_span);
    iy = rr.iy_lo + (r - q * rr.ny_span);
    iz = rr.iz_lo + q;
}

__device__ __forceinline__ RowRange row_range(const GridParams& g, float py, float pz, float R) {
    RowRange r;
    const float pad = R + 1e-3f * g.cell_yz;
    const int iy_lo = max(cell_coord(py - pad, g.oy, g.inv_cell_yz), 0);
    const int iy_hi = min(cell_coord(py + pad, g.oy, g.inv_cell_yz), g.ny - 1);
    const int iz_lo = max(cell_coord(pz - pad, g.oz, g.inv_cell_yz), 0);
    const int iz_hi = min(cell_coord(pz + pad, g.oz, g.inv_cell_yz), g.nz - 1);
    r.iy_lo = iy_lo;
    r.iz_lo = iz_lo;
    r.ny_span = iy_hi - iy_lo + 1;
    r.nrows = (iy_hi >= iy_lo && iz_hi >= iz_lo) ? r.ny_span * (iz_hi - iz_lo + 1) : 0;
    r.inv_span = 1.0f / (float)max(r.ny_span, 1);
    return r;
}

// point range [start,end) of row (iy,iz) that can hold points within R of (px,py,pz); conservative
__device__ __forceinline__ bool row_segment(const GridParams& g, const unsigned* __restrict__ cell_start, float px,
                                            float py, float pz, float R, int iy, int iz, unsigned& start,
                                            unsigned& end) {
    const float eps = 1e-3f * g.cell_yz;
    const float y0 = g.oy + (float)iy * g.cell_yz, z0 = g.oz + (float)iz * g.cell_yz;
    float dy = fmaxf(fmaxf(y0 - py, py - (y0 + g.cell_yz)), 0.0f);
    float dz = fmaxf(fmaxf(z0 - pz, pz - (z0 + g.cell_yz)), 0.0f);
    dy = fmaxf(dy - eps, 0.0f);
    dz = fmaxf(dz - eps, 0.0f);
    const float rem = R * R - dy * dy - dz * dz;
    if (rem < 0.0f) return false;
    const float xr = sqrtf(rem) + eps + R * 1e-6f;
    const int ix_lo = max(cell_coord(px - xr, g.ox, g.inv_cell), 0);
    const int ix_hi = min(cell_coord(px + xr, g.ox, g.inv_cell), g.nx - 1);
    if (ix_lo > ix_hi) return false;
    const unsigned row = ((unsigned)iz * g.ny + iy) * g.nx;
    start = __ldg(cell_start + row + ix_lo);
    end = __ldg(cell_start + row + ix_hi + 1);
    return end > start;
}

template <int MAXSEG, int MAXB = 2048>
struct SegList {
    unsigned start[MAXSEG];
    unsigned off[MAXSEG + 1];  // exclusive prefix of the lengths (len kept in off[] before the scan)
    unsigned short bseg[MAXB]; // segment that holds flattened candidate 32 * b (one search per 32 lanes)
    unsigned nseg;
    unsigned total;
};

// exclusive scan of the segment lengths (kept in off[]) + batch table. All NT threads call.
template <int NT, int MAXSEG, int MAXB, typename Sync>
__device__ __forceinline__ void finish_segments(SegList<MAXSEG, MAXB>& sl, unsigned tid, Sync&& group_sync) {
    if (tid < 32) {
        const unsigned n = sl.nseg;
        const unsigned per = (n + 31) / 32;
        const unsigned b = tid * per, e = min(n, b + per);
        unsigned s = 0;
        for (unsigned i = b; i < e; ++i) s += sl.off[i];
        unsigned inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= (unsigned)o) inc += up;
        }
        unsigned run = inc - s;
        for (unsigned i = b; i < e; ++i) {
            const unsigned len = sl.off[i];
            sl.off[i] = run;
            run += len;
        }
        if (tid == 31) { sl.off[n] = inc; sl.total = inc; }
    }
    group_sync();
    // batch table: segment of every 32nd candidate, so the per-candidate lookup is a short walk
    {
        const unsigned nb = min((sl.total + 31u) >> 5, (unsigned)MAXB);
        const unsigned nseg = sl.nseg;
        for (unsigned b = tid; b < nb; b += NT) {
            const unsigned j = b << 5;
            unsigned lo = 0, hi = nseg;
            while (hi - lo > 1) {
                const unsigned mid = (lo + hi) >> 1;
                if (sl.off[mid] <= j) lo = mid; else hi = mid;
            }
            sl.bseg[b] = (unsigned short)lo;
        }
    }
    group_sync();
}

// Enumerate the non-empty row segments of rows [row0, row0 + MAXSEG) of the query's row rectangle
// (start[] and the lengths in off[]); returns the number of candidate points. All NT threads call;
// NT == 32 only (the total is a warp reduction).
template <int NT, int MAXSEG, int MAXB, typename Sync>
__device__ __forceinline__ unsigned enumerate_segments(const GridParams& g, const unsigned* __restrict__ cell_start,
                                                       float px, float py, float pz, float R, const RowRange& rr, int row0,
                                                       SegList<MAXSEG, MAXB>& sl, unsigned tid, Sync&& group_sync) {
    static_assert(NT == 32, "enumerate_segments returns a warp-reduced total");
    if (tid == 0) sl.nseg = 0;
    group_sync();
    const int row1 = min(rr.nrows, row0 + MAXSEG);
    unsigned mine = 0;
    for (int r = row0 + (int)tid; r < row1; r += NT) {
        int iy, iz;
        row_coords(rr, r, iy, iz);
        unsigned s, e;
        if (row_segment(g, cell_start, px, py, pz, R, iy, iz, s, e)) {
            const unsigned slot = atomicAdd(&sl.nseg, 1u);
            sl.start[slot] = s;
            sl.off[slot] = e - s;
            mine += e - s;
        }
    }
    group_sync();
    return (unsigned)warp_sum((int)mine);
}

// Build the segment list for rows [row0, row0 + MAXSEG) of the query's row rectangle.
// All NT threads of the group call it. GROUP_SYNC: functor that synchronises the group.
template <int NT, int MAXSEG, int MAXB, typename Sync>
__device__ __forceinline__ void build_segments(const GridParams& g, const unsigned* __restrict__ cell_start, float px,
                                               float py, float pz, float R, const RowRange& rr, int row0,
                                               SegList<MAXSEG, MAXB>& sl, unsigned tid, Sync&& group_sync) {
    if (tid == 0) sl.nseg = 0;
    group_sync();
    const int row1 = min(rr.nrows, row0 + MAXSEG);
    for (int r = row0 + (int)tid; r < row1; r += NT) {
        int iy, iz;
        row_coords(rr, r, iy, iz);
        unsigned s, e;
        if (row_segment(g, cell_start, px, py, pz, R, iy, iz, s, e)) {
            const unsigned slot = atomicAdd(&sl.nseg, 1u);
            sl.start[slot] = s;
            sl.off[slot] = e - s;
        }
    }
    group_sync();
    finish_segments<NT>(sl, tid, group_sync);
}

// candidate j of the flattened list -> index into the cell-sorted array
template <int MAXSEG, int MAXB>
__device__ __forceinline__ unsigned seg_lookup(const SegList<MAXSEG, MAXB>& sl, unsigned j) {
    const unsigned b = j >> 5;
    unsigned s;
    if (b < (unsigned)MAXB) {
        s = sl.bseg[b];
        while (sl.off[s + 1] <= j) ++s;  // at most the segments that start inside this batch of 32
    } else {                             // beyond the table: plain binary search
        unsigned lo = 0, hi = sl.nseg;
        while (hi - lo > 1) {
            const unsigned mid = (lo + hi) >> 1;
            if (sl.off[mid] <= j) lo = mid; else hi = mid;
        }
        s = lo;
    }
    return sl.start[s] + (j - sl.off[s]);
}

}  // namespace bshot
