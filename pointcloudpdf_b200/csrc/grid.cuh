// grid.cuh -- layout of the neighbour-grid workspace shared by knn.cu (builder, query) and
// fps.cu (cell-ordered points for exact pruning).
#pragma once
#include "common.cuh"

namespace pob {

struct __align__(16) SceneGrid {
    float lox, loy, loz, inv_h;
    int dx, dy, dz, cell_base;
    float h;
    int start, end, use_grid;
    int pad0, pad1, pad2, pad3;
};
static_assert(sizeof(SceneGrid) == 64, "SceneGrid layout");

constexpr int MAX_DIM = 1024;          // cells per axis (bounds the rounding term in bound2)
constexpr int BRUTE_MAX_POINTS = 256;  // scenes this small are scanned directly
constexpr int SCAN_TILE = 2048;        // elements per scan block (512 threads x 4)

struct GridLayout {
    int64_t cap;  // capacity of the cell arrays
    size_t off_scene, off_bbox, off_cnt, off_start, off_tiles, off_pcell, off_sorted, off_fps, fps_bytes, total;
};

// scratch of the FPS launcher inside the grid workspace (fps.cu orders the points along a Hilbert curve before the
// cluster kernel reads them): ordered float4 copy | u64 keys x 2 | u32 values x 2 | radix-sort temporary
constexpr size_t FPS_SORT_TEMP_FIXED = 1 << 20;
static inline size_t fps_order_bytes(int64_t n) { return (size_t)n * (16 + 16 + 8 + 4) + FPS_SORT_TEMP_FIXED + 5 * 256; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline GridLayout grid_layout(int64_t n, int b, float cell_pts) {
    GridLayout L;
    if (!(cell_pts >= 0.25f)) cell_pts = 0.25f;
    L.cap = (int64_t)((double)n / (double)cell_pts) + b + 8;
    size_t o = 0;
    L.off_scene = o;  o = align_up(o + sizeof(SceneGrid) * (size_t)b, 256);
    L.off_bbox = o;   o = align_up(o + sizeof(int) * 6 * (size_t)b, 256);
    L.off_cnt = o;    o = align_up(o + sizeof(int) * (size_t)(L.cap + 1), 256);
    L.off_start = o;  o = align_up(o + sizeof(int) * (size_t)(L.cap + 1), 256);
    L.off_tiles = o;  o = align_up(o + sizeof(int) * (size_t)(L.cap / SCAN_TILE + 2), 256);
    L.off_pcell = o;  o = align_up(o + sizeof(int) * (size_t)n, 256);
    L.off_sorted = o; o = align_up(o + sizeof(float4) * (size_t)n, 256);
    L.off_fps = o;    L.fps_bytes = fps_order_bytes(n); o = align_up(o + L.fps_bytes, 256);
    L.total = o;
    return L;
}


}  // namespace pob
