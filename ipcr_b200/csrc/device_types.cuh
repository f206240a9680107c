// device_types.cuh -- structures shared by the host orchestration and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ipcr_gpu.h"

namespace ipcr {

// ---- packed genome layout ("striped bit-planes") ----------------------------
// The concatenated ASCII buffer of a genome (all segments back to back) is cut
// into STRIPES of W consecutive positions.  Stripe g = P / W, row w = P % W.
// 1024 consecutive stripes form a TILE: lane l = (g / 32) % 32, bit i = g % 32.
// A tile is stored as (W + 31) ROWS of 32 lanes x 16 bytes; the uint4 at
// (tile, row r, lane l) holds four one-hot plane words {A, C, G, T} whose bit i
// describes position ((tile*1024 + l*32 + i) * W + r).  Rows W..W+30 repeat the
// first 31 positions of the following stripes so that a window of up to 32
// consecutive positions starting at any row < W is rows r..r+31 of the SAME
// lane and bit: no shifts, no cross-lane traffic.  Every byte that is not an
// uppercase A/C/G/T (N, IUPAC codes, lowercase, padding) has all four bits
// clear, i.e. it is a hard mismatch (core/primer/iupac.go:62-67).
// 4 bits per base: 0.5 B/bp (+31/W for the repeated rows).
constexpr int kLanes = 32;
constexpr int kStripesPerTile = 1024;
constexpr int kExtRows = 31;

struct GenomeLayout {
    uint64_t total_len;  // bases in the concatenated buffer
    uint32_t W;          // rows per stripe
    uint32_t rows;       // W + kExtRows
    uint64_t n_tiles;
};

struct SegmentDev {
    uint64_t global_off;  // offset in the concatenated buffer
    uint64_t len;
    uint64_t own_len;
    uint64_t rec_off;
    uint32_t rec_idx;
    uint32_t pad;
};

// ---- panel tables on the device ---------------------------------------------
struct OrientDev {
    uint16_t L;
    uint16_t seed_off;
    uint16_t seed_len;
    uint8_t rc_type;
    uint8_t seeded;
    uint32_t cls_off;    // offset into the class byte array
    uint16_t j0[2];      // filter window offset per plan variant (candidate start = window start - j0)

    uint64_t prot_mask;  // protected columns (bit j), empty when terminal window is off
};

// table-driven filter program of one orientation (generic kernel)
struct OrientProg {
    uint32_t col_begin;   // into cols[]
    uint32_t blk_begin;   // into blk_len[]
    uint16_t n_prot;
    uint16_t n_blocks;
    uint16_t need;
    uint16_t j0;
};

struct ColDev {
    uint8_t row;
    uint8_t slot;  // index into the used-class ring
};

struct PairDev {
    uint16_t fwdA, fwdB, rcA, rcB;
    uint16_t alen, blen;
    uint32_t pad;
    int64_t min_len, max_len;
};

struct ScanParams {
    int32_t max_mm;
    int32_t terminal_window;
    int64_t hit_cap;
    uint8_t circular;
    uint8_t raw_rc;  // plan variant in use
    uint8_t pad[6];
};

// candidate word: orient << 44 | global start in the concatenated buffer
constexpr int kCandPosBits = 44;
constexpr uint64_t kCandPosMask = (1ull << kCandPosBits) - 1;

// sorted-site record used by the pairing kernels
struct FinalSite {
    uint64_t pos;
    uint64_t mask;
    uint32_t mm;
    uint32_t pad;
};

}  // namespace ipcr
