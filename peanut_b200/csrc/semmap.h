// Stage B (Semantic_Mapping) device state: see semmap.cu.
#pragma once
#include "engine.h"

namespace pn {

// Geometry constants, precomputed on the host the way Semantic_Mapping.__init__ does (mapping.py:12-50).
struct SemMapCfg {
  int h, w;            // mapper frame (120 x 160)
  int channels;        // 4 + num_sem_categories
  int nf;              // 1 + num_sem_categories (count channel + semantic features)
  int ego_channels;    // 2 + num_sem_categories (obstacle, explored, categories)
  int vr;              // vision_range (cells of the ego window)
  int nz;              // height bins
  int min_z, max_z;    // agent-height slab [min_z, max_z)
  int map_cells;       // local map side
  int special_f[3];    // feature indices projected over all heights (mapping.py:106-113), -1 = unused
  float xc, zc, f;     // camera matrix (depth_utils.py:27-34)
  float agent_height, shift_x, res, half_vr, vr_f, z_mid, nz_f;
  float map_thr, exp_thr, cat_thr;
  float fuse_r2 = 0.f;  // set by SemMap::init: squared reach of the ego window in k_fuse's sampling space
};

struct SemMap {
  SemMapCfg c{};
  int E = 0;
  Arena arena;
  float* coords = nullptr;     // [E][3][N] normalised splat coordinates
  int* col_count = nullptr;    // [E][vr*vr]
  uint32_t* qcount = nullptr;  // [E][4] valid heights / heights in the stair band / heights <= 0.2 (same allocation as col_count)
  int* col_start = nullptr;    // [E][vr*vr + 1]
  int* col_fill = nullptr;     // [E][vr*vr]
  uint32_t* entries = nullptr; // [E][4*N] bucketed (corner, z, point) keys
  float* ego = nullptr;        // [E][ego_channels][vr][vr]
  int* col_list = nullptr;     // [E][2][vr*vr] non-empty columns: plane 0 = tiny from the front, big from the back; plane 1 = mid / large
  int* list_n = nullptr;       // [E][4] = {#tiny, #big, #mid, #large}
  int num_sms = 148;
  float* xf = nullptr;         // [E][4] cos, sin, tx, ty of the sampling grids
  int* stair_flag = nullptr;   // [E]
  static constexpr int kLaunches = 11;  // kernels per forward (plus two memsets)

  void init(const SemMapCfg& cfg, int envs);
  // maps_last may be a strided view (element strides between envs, channel planes and rows; unit x stride)
  void forward(const float* obs, const float* pose_delta, const float* maps_last, long long ml_env, long long ml_plane,
               long long ml_row, float* poses_inout, float* fp_out, float* map_out, cudaStream_t s);
};

}  // namespace pn
