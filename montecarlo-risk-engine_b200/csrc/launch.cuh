// Host-side helpers shared by the entry points: path-shard descriptor, RNG descriptor
// conversion and validation (include/mcre.h: mcre_rng, mcre_shard).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace mcre {

struct ShardDev {
  long long path_begin, n_paths;
  int chunk;
};

inline RngDev make_rng(const mcre_rng *r) {
  RngDev d;
  d.mode = r->mode; d.k0 = (uint32_t)r->seed; d.k1 = (uint32_t)r->stream;
  d.z = r->d_z; d.u = r->d_u; d.n_total = r->n_paths_total;
  return d;
}
inline int check_shard(const mcre_shard *s) {
  if (!s || s->n_paths < 0 || s->chunk_paths <= 0 || s->chunk_paths % 256 != 0)
    return fail(-2, "invalid shard: chunk_paths must be a positive multiple of 256%s", "");
  if (s->path_begin % s->chunk_paths != 0) return fail(-2, "invalid shard: path_begin not chunk aligned%s", "");
  return 0;
}
inline double pack2(int lo, int hi) {
  long long v = ((long long)(unsigned int)hi << 32) | (unsigned int)lo;
  double d;
  memcpy(&d, &v, 8);
  return d;
}

}  // namespace mcre
