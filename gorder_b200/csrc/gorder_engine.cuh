// gorder_engine.cuh — device-visible data model of the engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gorder_b200.h"

namespace gorder {

constexpr int kBlock = 256;        // threads per CTA of the accumulation kernels
constexpr int kWarps = kBlock / 32;
constexpr int kMolAlign = 32;      // alignment of the extra (non-molecule) planes, in floats

// One molecule type on the device.  Coordinates of a frame live in TILED "planes": molecules are
// grouped in tiles of `tile` consecutive molecules (= the molecules one CTA of the accumulation
// kernel owns); inside a tile, for used atom u and component c, molecule m sits at
//     frame[plane_base + (m / tile) * tile_stride + c * cstride + u * tile + (m % tile)],
//     cstride = n_used * tile,  tile_stride = 3 * cstride
// (component-major inside the tile: the leaflet-axis planes of a tile form one contiguous run for the
// centre reduction).
// A warp whose lanes are consecutive molecules reads fully-used 128 B lines, and ONE CTA streams ONE
// contiguous region of tile_stride floats (S-CG: 147 KB) -- the DRAM-friendly pattern of a plain copy.
struct TypeDesc {
    int n_mol;        // molecules of this type
    int mpad;         // n_mol rounded up to a whole number of tiles
    int tile;         // molecules per tile
    int tile_stride;  // floats per tile
    int cstride;      // component stride inside a tile (n_used * tile)
    int n_orders;     // order slots (bond types, or virtual C-H bonds for UA)
    int n_items;      // bond types (AA/CG) or carbon types (UA)
    int plane_base;   // float offset of plane (u=0, c=0) inside a frame
    int slot0;        // first order slot of this type
    int molpad0;      // offset of this type in padded per-molecule arrays (leaflets, normals)
    int mol0;         // offset of this type in compact per-molecule arrays (exports)
    int item_off;     // offset of this type in the item tables
    int head_off;     // in-tile offset (u * tile) of the leaflet head, -1 if none
    int nhead_off;    // plane offset of the dynamic-normal head, -1 if none
    int n_methyls;
    int methyl_off;   // offset into methyl plane-offset table
    int manual_leaf_off;   // byte offset of this type's manual leaflet rows, -1 if none
    int n_manual_leaf;     // rows available
    int manual_norm_off;   // float offset of this type's manual normals, -1 if none
    int n_manual_norm;
};

// AA/CG bond type: in-tile offsets (u * tile) of the two atoms' x planes.
struct BondItem { int a_off, b_off; };

// UA carbon type (uaorder.rs:234-239): plane offsets of target and helpers, kind, first slot.
struct UAItem { int kind, t_off, h1_off, h2_off, h3_off, slot_rel, pad0, pad1; };

struct Chunk { int type, first_mol; };

// Per-frame data of a batch.
struct FrameAux {
    float L[3];           // box (0 if !handle_pbc)
    float half[3];        // L / 2
    long long frame_index;
    int leaf_row;         // row of the leaflet table this frame uses (-1: none)
    int tw_row;           // row of the per-frame accumulators
    int manual_norm_row;  // frame_index / step
    int pad;
    // geometry shape of this frame (geometry.rs construct_shape)
    float shape_origin[3];
    float shape_len[3];   // cuboid extents (INFINITY allowed)
    float shape_radius, shape_height;
    float center[3];      // scratch: group centre (geometry reference)
    float guard[3];       // 0.99 L / 2: below it the minimum-image fold takes its exact fast path
};

// Group of atoms given by native offsets (membrane, geometry reference, normal heads).
struct GroupRef {
    const int *off;   // native offset of x
    const int *cs;    // component stride
    const int *slot;  // original slot (for error reports)
    int n;
};

// Run of contiguous floats inside a native frame (one component of a group's atoms).
struct Seg { int off, len; };

struct MapParams {
    int enabled, plane, nx, ny;
    float x0, y0, binx, biny;
    float inv_binx, inv_biny;   // 1 / bin: the bin look-up divides only next to a bin edge
    long long n_bins;
};

struct ShapeParams {
    int kind, invert, axis, ref_kind;
    float ref_point[3];
    float dims[6];
    float structure_box[3];   // POINT reference: the shape's origin is wrapped with the structure file's box (0: the frame's)
};

// Everything the kernels need, passed by value.
struct DeviceView {
    const TypeDesc *types;
    const Chunk *chunks;
    const BondItem *bonds;
    const UAItem *ua;
    const int *methyl_offs;
    int n_types, n_chunks, n_slots, n_molpad, n_mol_total;
    long long frame_floats;   // floats per native frame
    // configuration
    int kind, handle_pbc, step;
    int normal_mode, normal_axis, collect_normals;
    float dynamic_radius;
    int leaflet_mode, leaflet_axis, leaflet_flip, leaflet_freq_kind, leaflet_freq;
    float leaflet_radius;
    ShapeParams shape;
    MapParams map;
    GroupRef membrane, geom_ref, normal_heads;
    const unsigned char *manual_leaflets;
    const float *manual_normals;
    // UA rotation constants (sin, cos) computed on the host with the same libm as the reference
    float tet_s, tet_c, tet_half_s, tet_half_c, ch3_s, ch3_c;
    int ua_exact;             // UA: bit-exact hydrogen construction everywhere (default: only with geometry / maps)
    // error word
    int *err;                 // [0] code, [1] unused
    long long *err_detail;
};

}  // namespace gorder
