// mdb_dist.cuh -- per-rank device steps of the 1-D block-cyclic multi-GPU factorization.
// Included at the end of mdb_api.cu.
#pragma once

struct mdb_dist {
  mdb_ctx* ctx;
};

extern "C" int mdb_dist_create(mdb_ctx* ctx, int64_t, int, int, mdb_dist** out) {
  if (out) *out = nullptr;
  return mdb_fail(ctx, MDB_ERR_STATE, "%s", "mdb_dist: not built yet");
}
extern "C" int mdb_dist_destroy(mdb_dist*) { return MDB_OK; }
extern "C" int mdb_dist_generate_f2(mdb_dist*, uint64_t, double, const int64_t*) { return MDB_ERR_STATE; }
extern "C" int mdb_dist_set_matrix(mdb_dist*, const double*, int64_t) { return MDB_ERR_STATE; }
extern "C" int mdb_dist_set_bounds(mdb_dist*, const mdb_bounds*, const int64_t*) { return MDB_ERR_STATE; }
extern "C" int64_t mdb_dist_panel_elems(const mdb_dist*, int64_t) { return 0; }
extern "C" void* mdb_dist_panel_buffer(mdb_dist*, int) { return nullptr; }
extern "C" int mdb_dist_factor_panel(mdb_dist*, int64_t, int) { return MDB_ERR_STATE; }
extern "C" int mdb_dist_apply_panel(mdb_dist*, int64_t, int, int64_t, int64_t) { return MDB_ERR_STATE; }
extern "C" int mdb_dist_get_local(mdb_dist*, double*, int64_t, int, double*, double*, uint8_t*) { return MDB_ERR_STATE; }
extern "C" int64_t mdb_dist_local_cols(const mdb_dist*) { return 0; }
