// Blob layout shared by host and device code; mirrors tsim_b200/pack.py.
#pragma once
#include <stdint.h>

namespace tsb {

constexpr uint32_t kMagic = 0x32425354u;  // "TSB2"
constexpr uint32_t kVersion = 8;
constexpr int kModeFaithful = 0;
constexpr int kModeFast = 1;
constexpr int kModeSliced = 2;

constexpr int kHeaderWords = 32;
constexpr int kCompWords = 8;
constexpr int kLevelWords = 12;
constexpr int kChunkWords = 4;
constexpr int kPrefactorWords = 8;

enum HeaderSlot {
  H_MAGIC = 0, H_VERSION, H_MODE, H_W,
  H_NUM_F, H_N_OUT, H_N_DIRECT, H_N_COMP,
  H_N_DRAWS, H_N_LEVELS, H_N_CHUNKS, H_MAX_CHUNK,
  H_OFF_DIRECT, H_OFF_COMP, H_OFF_LEVEL, H_OFF_CHUNK,
  H_OFF_FSEL, H_OFF_DEST, H_OFF_DATA, H_DATA_WORDS,
  H_TOTAL_WORDS, H_WF64, H_WOUT64, H_OFF_TABLES,
  H_TABLE_WORDS, H_ONE_ROW, H_ZERO_ROW, H_PLANE_ROWS,
  H_INDEX_SCALE  // sliced records: row index bytes are stored times this (2 when rows <= 127, so that 64-bit lanes can address them)
};

// component table row
enum CompSlot { C_F = 0, C_NC, C_FSEL_OFF, C_FIRST_DRAW, C_FIRST_LEVEL, C_N_LEVELS };
// level table row
enum LevelSlot { L_G = 0, L_P, L_A, L_H, L_C, L_D, L_FLAGS, L_FIRST_CHUNK, L_N_CHUNKS, L_P_LO };
// chunk table row
enum ChunkSlot { K_OFF = 0, K_WORDS, K_GRAPHS };

}  // namespace tsb
